#pragma once
#include "bc_mocks.hpp"
