// TEST INFRASTRUCTURE ONLY.  Named by the reference headers (Frame.h, MapPlane.h ...) for point-cloud members the matcher / optimiser paths never touch.
#pragma once
#include <pcl/point_cloud.h>
