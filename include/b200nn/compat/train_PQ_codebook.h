// Forwarding header: `#include "train_PQ_codebook.h"` (opq/train_codebook/train_PQ.cpp:1) -> the GPU trainer.
#pragma once
#include <iostream>
#include <string>
#include "../train_pq_gpu.hpp"
using namespace std;  // the reference header injects it (train_PQ_codebook.h:21) and its main relies on that
