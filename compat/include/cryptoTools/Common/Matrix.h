#pragma once
// compat: cryptoTools/Common/Matrix.h
#include "aby3_b200/sh3/Defines.h"
#include "compat_surface.h"
