#pragma once
// compat: <aby3/Common/Defines.h> (reference: aby3/Common/Defines.h:1-70)
#include "aby3_b200/sh3/Defines.h"
#include "compat_surface.h"
