#pragma once
// compat: the SharedOT class lives with the facade's evaluator (aby3/OT/SharedOT.h:1-49)
#include "aby3_b200/sh3/Sh3Evaluator.h"
