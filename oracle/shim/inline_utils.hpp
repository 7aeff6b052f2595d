// oracle/shim/inline_utils.hpp — TEST INFRASTRUCTURE ONLY.
// Shadows the reference's <inline_utils.hpp> (it is included with <>, core_private.cpp:4): pulls in
// the real header unchanged, then redirects the two mtrand call sites of
// opt_guess_translational_motion (core_private.cpp:42-43; `i` is that function's loop variable)
// to the pinned counter-based RNG.
#pragma once
#include_next <inline_utils.hpp>
#include "ref_context.hpp"
#define mtrand(lo, hi) rssync_ref::pinned_mtrand(i, __LINE__, (lo), (hi))
