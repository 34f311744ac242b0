// TEST INFRASTRUCTURE.  apps/demo.cpp includes this header but calls nothing from it (its cv::normalize line is commented out).
#pragma once
