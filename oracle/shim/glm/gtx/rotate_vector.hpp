// TEST INFRASTRUCTURE ONLY: nothing from gtx/rotate_vector is used on the hot path.
#pragma once
#include "../glm.hpp"
