// Shadows localization_interface/localization_struct.hpp for the oracle/_ref build: the EKF / GNSS structs it declares are
// not used by registration.{hpp,cpp} / voxel_hash_map.{hpp,cpp}, and they need Eigen::Quaterniond, which the stand-in lacks.
#pragma once
#include "Eigen/Dense"
