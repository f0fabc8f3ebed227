// TEST INFRASTRUCTURE ONLY: glm::rotate lives in ../glm.hpp of this shim.
#pragma once
#include "../glm.hpp"
