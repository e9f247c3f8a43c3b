#pragma once
// compat: cryptoTools/Crypto/AES.h -> aby3_b200/sh3/Crypto.h
#include "aby3_b200/sh3/Crypto.h"
#include "compat_surface.h"
