// bbfft/user_module.hpp -- same include name as the reference; everything lives in bbfft/api.hpp.
#ifndef BBFFT_FWD_USER_MODULE_HPP
#define BBFFT_FWD_USER_MODULE_HPP
#include "bbfft/api.hpp"
#endif
