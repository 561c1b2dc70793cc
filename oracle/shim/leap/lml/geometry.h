// TEST INFRASTRUCTURE — stand-in for leap/lml/geometry.h (see vector.h in this directory).
#pragma once
#include "vector.h"
