// drop-in header: same path as in dgrzech/sobfu; the declarations live in sobfu_b200_shim.hpp
#pragma once
#include <sobfu_b200_shim.hpp>
