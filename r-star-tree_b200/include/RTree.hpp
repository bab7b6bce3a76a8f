// RTree.hpp -- umbrella header of the rtree-b200 host library (header-only C++17).
//
//   #include <RTree.hpp>            // namespace eh::rtree, same API as ehwan/R-Star-Tree
//   eh::rtree::RTree<box, key, mapped> tree;  tree.insert(...);
//   auto image = tree.flatten_device();       // new: GPU hand-off (device_layout.hpp)
//   rt_tree* gpu; rt_upload(&image.desc(), 0, &gpu);   // include/rtree_b200.h
#pragma once

#include "RTree/device_layout.hpp"
#include "RTree/geometry.hpp"
#include "RTree/node.hpp"
#include "RTree/rtree.hpp"
#include "RTree/split.hpp"
