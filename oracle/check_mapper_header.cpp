// check_mapper_header.cpp -- compile-only check (oracle/build_ref.py, where /root/reference exists) that
// include/sage_ba_mapper.hpp fits the reference's OWN map types: df::Map<float> / df::Keyframe<float> / df::Frame<float>
// (core/mapping/keyframe_map.h:93-120, keyframe.h:19-61, frame.h:16-125) with the Sophus the reference vendors.  The explicit
// instantiation forces every member function to compile; nothing is linked or run.
#include "keyframe_map.h"

#include "sage_ba_mapper.hpp"

template class sage::BatchedLocalBA<df::Map<float>, 16>;
template class sage::BatchedLocalBA<df::Map<float>, 32>;

int main() { return 0; }
