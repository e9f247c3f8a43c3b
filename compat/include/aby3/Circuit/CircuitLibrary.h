#pragma once
// compat: aby3::CircuitLibrary (aby3/Circuit/CircuitLibrary.h:10-52) = oc::BetaLibrary + the aby3-specific builders;
// the facade's library already carries those (int_Sh3Piecewise_helper, int_comp_helper, bits_nor_helper).
#include "aby3_b200/sh3/BetaCircuit.h"
namespace aby3 {
class CircuitLibrary : public oc::BetaLibrary {};
}  // namespace aby3
