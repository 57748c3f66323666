// bbfft/detail/plan_impl.hpp -- same include name as the reference; see bbfft/api.hpp.
#ifndef BBFFT_FWD_DETAIL_PLAN_IMPL_HPP
#define BBFFT_FWD_DETAIL_PLAN_IMPL_HPP
#include "bbfft/api.hpp"
#endif
