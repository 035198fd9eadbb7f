// Stand-in for <ceres/rotation.h> — see ceres.h in this directory.
#pragma once
#include "ceres.h"
