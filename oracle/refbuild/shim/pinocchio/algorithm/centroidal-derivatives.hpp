// TEST INFRASTRUCTURE ONLY: forwards to the Pinocchio stand-in (see cafe_pinocchio_shim.hpp)
#pragma once
#include "../cafe_pinocchio_shim.hpp"
