// bbfft/plan.hpp -- same include name as the reference; everything lives in bbfft/api.hpp.
#ifndef BBFFT_FWD_PLAN_HPP
#define BBFFT_FWD_PLAN_HPP
#include "bbfft/api.hpp"
#endif
