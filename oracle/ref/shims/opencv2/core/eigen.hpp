// TEST INFRASTRUCTURE ONLY - forwards to the stand-in of oracle/ref/shims/opencv2/opencv.hpp (cv2eigen / eigen2cv are declared there when needed)
#pragma once
#include <opencv2/opencv.hpp>
#include <Eigen/Core>
