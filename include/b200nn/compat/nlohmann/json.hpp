// Placeholder for <nlohmann/json.hpp>: scalar_quantization/scalar_quantization/int8_quan_test.cpp:6 includes it but only
// uses it inside a commented-out block, and the drop-in Int8Quan (b200nn/int8_quan_gpu.hpp) scans its tiny model-list
// JSON by hand.  Nothing is declared here on purpose.
#pragma once
