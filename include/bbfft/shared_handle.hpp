// bbfft/shared_handle.hpp -- same include name as the reference; everything lives in bbfft/api.hpp.
#ifndef BBFFT_FWD_SHARED_HANDLE_HPP
#define BBFFT_FWD_SHARED_HANDLE_HPP
#include "bbfft/api.hpp"
#endif
