// ORACLE / TEST INFRASTRUCTURE: stand-in for <gtsam/base/numericalDerivative.h>, see ../mini_gtsam.h
#pragma once
#include "../mini_gtsam.h"
