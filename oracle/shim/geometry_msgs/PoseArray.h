// ORACLE / TEST INFRASTRUCTURE: see PoseStamped.h
#pragma once
#include "PoseStamped.h"
