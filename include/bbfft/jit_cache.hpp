// bbfft/jit_cache.hpp -- same include name as the reference; everything lives in bbfft/api.hpp.
#ifndef BBFFT_FWD_JIT_CACHE_HPP
#define BBFFT_FWD_JIT_CACHE_HPP
#include "bbfft/api.hpp"
#endif
