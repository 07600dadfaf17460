// ORACLE / TEST INFRASTRUCTURE: stand-in for <gtsam/geometry/Pose3.h>, see ../mini_gtsam.h
#pragma once
#include "../mini_gtsam.h"
