#pragma once
// compat: cryptoTools/Network/Channel.h -> the facade's channel + the in-process session shim
#include "compat_net.h"
