// TEST INFRASTRUCTURE ONLY: see array2d.h
#pragma once
#include "array2d.h"
