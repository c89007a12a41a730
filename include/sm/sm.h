// sm/sm.h -- umbrella header, drop-in for the reference's include/sm.h:2-3.
// Build user code with  -I<repo>/include -I<repo>/include/sm  and link
// -L<repo>/simplemath_b200 -lsmb200.
#pragma once
#include "SMArray.h"
#include "UserFunctions.h"
#include "Lazy.h"    // opt-in op-chain fusion (not in the reference)
#include "Runtime.h" // opt-in async scope / device set (not in the reference)
