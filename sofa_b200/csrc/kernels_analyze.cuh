#pragma once
#include "kernels_emit.cuh"
namespace sofa {}
