#pragma once
// shim of boost/multiprecision/cpp_int.hpp: int128_t (Sh3FixedPoint.cpp:11)
namespace boost { namespace multiprecision { typedef __int128 int128_t; typedef unsigned __int128 uint128_t; } }
