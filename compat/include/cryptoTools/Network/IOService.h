#pragma once
// compat: cryptoTools/Network/IOService.h -> the facade's channel + the in-process session shim
#include "compat_net.h"
