// TEST INFRASTRUCTURE: see opencv.hpp in this directory.
#pragma once
#include "opencv.hpp"
