// Forwarding stand-in: see node_world.hpp.
#pragma once
#include "../node_world.hpp"
