// bbfft/configuration.hpp -- same include name as the reference; everything lives in bbfft/api.hpp.
#ifndef BBFFT_FWD_CONFIGURATION_HPP
#define BBFFT_FWD_CONFIGURATION_HPP
#include "bbfft/api.hpp"
#endif
