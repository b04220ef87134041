// Forwarding header: `#include "int8_quan.h"` (scalar_quantization/scalar_quantization/int8_quan_test.cpp:8).
#pragma once
#include "../int8_quan_gpu.hpp"
