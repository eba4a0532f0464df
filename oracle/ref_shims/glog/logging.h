// Stub for <glog/logging.h>, reached through the reference's common/indexed_map.h (LOG(FATAL) / CHECK in IndexedMap::Get) when
// oracle/check_mapper_header.cpp includes core/mapping/keyframe_map.h: c10's glog-compatible macros stand in.
#pragma once
#include <c10/util/Logging.h>
