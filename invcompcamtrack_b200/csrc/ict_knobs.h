// ict_knobs.h — A/B switches of profiling runs.
//
// The default build has NO environment-dependent behaviour: ict_knob() is a constant null pointer, so every
// `if (ict_knob("..."))` below folds away and an exported variable can never change which kernel runs or what it
// computes.  A profiling build (python -m invcompcamtrack_b200.build with ICT_EXTRA_NVCC=-DICT_PROFILING) reads the
// variables; the tools under profiles/tools/ that sweep variants say so in their headers.  The two switches tests and
// tools need in the default build are explicit state of the tracker instead: ict_tracker_set_knob (ictrack.h).
#pragma once
#include <stdlib.h>

#ifdef ICT_PROFILING
static inline const char* ict_knob(const char* name) { return getenv(name); }
#else
static inline const char* ict_knob(const char*) { return nullptr; }
#endif
