// bbfft/jit_cache_all.hpp -- same include name as the reference; everything lives in bbfft/api.hpp.
#ifndef BBFFT_FWD_JIT_CACHE_ALL_HPP
#define BBFFT_FWD_JIT_CACHE_ALL_HPP
#include "bbfft/api.hpp"
#endif
