// TEST INFRASTRUCTURE ONLY (oracle/_ref build): stand-in for HydraAPI's aligned_alloc.h (not vendored).
#pragma once
#include <vector>
namespace cvex { template<class T> using vector = std::vector<T>; }
