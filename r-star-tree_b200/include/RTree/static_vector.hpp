// RTree/static_vector.hpp -- kept so that reference-style includes keep working.
#pragma once
#include "node.hpp"
