// bbfft/module_format.hpp -- same include name as the reference; everything lives in bbfft/api.hpp.
#ifndef BBFFT_FWD_MODULE_FORMAT_HPP
#define BBFFT_FWD_MODULE_FORMAT_HPP
#include "bbfft/api.hpp"
#endif
