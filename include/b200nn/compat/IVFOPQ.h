// Forwarding header: `#include "IVFOPQ.h"` (opq/src/multi_frame_index_test.cpp:3) -> the GPU class.
#pragma once
#include "../ivfopq_gpu.hpp"
using namespace std;  // the reference header injects it (IVFOPQ.h:17) and its callers rely on that
