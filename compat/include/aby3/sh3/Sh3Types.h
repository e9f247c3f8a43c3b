#pragma once
// compat: <aby3/sh3/Sh3Types.h> of the reference tree -> the B200 facade's header of the same name
#include "aby3_b200/sh3/Sh3Types.h"
#include "compat_surface.h"
