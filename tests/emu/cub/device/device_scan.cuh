// TEST INFRASTRUCTURE ONLY: see ../cub.cuh
#pragma once
#include "../cub.cuh"
