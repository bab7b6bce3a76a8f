// RTree/rstar_split.hpp -- kept so that reference-style includes keep working.
#pragma once
#include "split.hpp"
