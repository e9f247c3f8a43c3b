#pragma once
// compat: cryptoTools/Circuit/BetaCircuit.h -> aby3_b200/sh3/BetaCircuit.h
#include "aby3_b200/sh3/BetaCircuit.h"
