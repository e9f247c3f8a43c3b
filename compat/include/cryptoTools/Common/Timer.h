#pragma once
// compat: cryptoTools/Common/Timer.h
#include "aby3_b200/sh3/Defines.h"
#include "compat_surface.h"
