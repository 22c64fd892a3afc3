// Empty stand-in: registration.{hpp,cpp} and voxel_hash_map.{hpp,cpp} include this PCL header but use nothing from it.
#pragma once
