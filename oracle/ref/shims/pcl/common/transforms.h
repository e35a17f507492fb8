// TEST INFRASTRUCTURE ONLY - forwards to the compile-only stand-in of oracle/ref/shims/pcl/point_cloud.h
#pragma once
#include <pcl/point_cloud.h>
