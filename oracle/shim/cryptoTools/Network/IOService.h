#pragma once
// shim of cryptoTools/Network/IOService.h: the IOService class lives with the Session stand-in
#include "cryptoTools/Network/Session.h"
