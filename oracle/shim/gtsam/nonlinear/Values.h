// ORACLE / TEST INFRASTRUCTURE: stand-in for <gtsam/nonlinear/Values.h>, see ../mini_gtsam_lm.h
#pragma once
#include "../mini_gtsam_lm.h"
