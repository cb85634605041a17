// TEST INFRASTRUCTURE ONLY: see accumulators.hpp in this directory.
#pragma once
#include "boost/accumulators/accumulators.hpp"
