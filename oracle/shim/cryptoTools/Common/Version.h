#pragma once
// shim: the reference only checks CRYPTO_TOOLS_VERSION >= 10601 (aby3/Common/Defines.h:5-7)
#define CRYPTO_TOOLS_VERSION 10601
