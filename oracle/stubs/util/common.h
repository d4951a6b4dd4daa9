// Stand-in for the reference's util/common.h (see README.md): the two helpers matching_io.cpp
// calls.  Their definitions are the reference's own (src/util/common.cpp:40-48, 85-135), which the
// Makefile compiles from that file by line range.
#pragma once
#include <iostream>
#include <set>
#include <string>
#include <vector>
#include <data_structures/track.h>
namespace orthosfm {
std::string zfill(const int& value, const int& zeros);
std::vector<Track> filterTracksToAvailableCameras(const std::vector<unsigned int>& ids, const std::vector<Track>& tracks,
                                                  bool onlyFullSizeTracks, bool keepAdditionalCamera);
}  // namespace orthosfm
