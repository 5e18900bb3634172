#pragma once
#include "fedd_mocks.hpp"
