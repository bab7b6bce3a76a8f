// RTree/static_node.hpp -- kept so that reference-style includes keep working.
#pragma once
#include "node.hpp"
