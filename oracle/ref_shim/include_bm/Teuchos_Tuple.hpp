#pragma once
#include "bm_mocks.hpp"
