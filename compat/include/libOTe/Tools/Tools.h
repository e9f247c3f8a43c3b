#pragma once
// compat: libOTe/Tools/Tools.h (oc::transpose) -- the facade transposes on the device
#include "aby3_b200/sh3/Defines.h"
