#pragma once
// compat: cryptoTools/Circuit/BetaLibrary.h -> aby3_b200/sh3/BetaCircuit.h
#include "aby3_b200/sh3/BetaCircuit.h"
