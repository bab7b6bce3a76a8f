// RTree/geometry_traits.hpp -- kept so that reference-style includes keep working.
#pragma once
#include "geometry.hpp"
