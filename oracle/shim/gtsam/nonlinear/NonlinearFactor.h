// ORACLE / TEST INFRASTRUCTURE: stand-in for <gtsam/nonlinear/NonlinearFactor.h>, see ../mini_gtsam.h
#pragma once
#include "../mini_gtsam.h"
