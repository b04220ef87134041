// Forwarding header: lets the reference's own sources (#include "brutoforce.h") compile UNMODIFIED against the
// B200 index -- add -I include/b200nn/compat in front of the reference's include path.
#pragma once
#include "../hnswlib_gpu.hpp"
