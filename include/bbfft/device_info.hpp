// bbfft/device_info.hpp -- same include name as the reference; everything lives in bbfft/api.hpp.
#ifndef BBFFT_FWD_DEVICE_INFO_HPP
#define BBFFT_FWD_DEVICE_INFO_HPP
#include "bbfft/api.hpp"
#endif
