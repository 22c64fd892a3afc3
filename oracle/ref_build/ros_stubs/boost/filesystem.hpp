// Empty stand-in (ROS / tf / OpenCV / boost are absent from this image): the EKF sources include this header; what they use from it,
// if anything, is declared in ros/ros.h of this directory.
#pragma once
#include "ros/ros.h"
