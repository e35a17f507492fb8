// TEST INFRASTRUCTURE ONLY - forwards to the stand-in of oracle/ref/shims/opencv2/opencv.hpp
#pragma once
#include <opencv2/opencv.hpp>
