// TEST INFRASTRUCTURE ONLY - compile-only stand-in, see oracle/ref/shims/pcl/point_cloud.h
#pragma once
#include <pcl/point_cloud.h>
