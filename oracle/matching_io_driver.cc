/*
 * matching_io_driver.cc -- extern "C" driver around the reference's own track-file code.
 * TEST INFRASTRUCTURE ONLY.
 *
 * Linked (oracle/Makefile, _ref/libmatching_io_ref.so) with the UNMODIFIED
 * /root/reference/src/matching/matching_io.cpp (saveTracksToFile, loadTracksFromFile,
 * saveTracksToPairwiseFiles: lines 16-140), src/data_structures/track.cpp and the two helpers of
 * src/util/common.cpp the writer calls, compiled against the stand-in third-party headers in
 * oracle/stubs/.  tests/golden/make_golden.py uses it to write the golden tracks.txt /
 * AAA_BBB.txt files; tests hold osfm_io_save_tracks / osfm_io_save_pairwise_tracks /
 * osfm_io_load_tracks (orthosfm_b200/csrc/io_formats.cuh) against it byte for byte.
 *
 * orthosfm::View is only asked for its id by saveTracksToPairwiseFiles; its constructor and
 * getID() are defined here (src/data_structures/view.cpp needs OpenCV's image I/O).
 */
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include <matching/matching_io.h>

orthosfm::View::View(const unsigned int& id, const std::string& imagePath) : m_id(id), m_imagePath(imagePath)
{
    std::memset(&m_siftData, 0, sizeof m_siftData);
}
const unsigned int& orthosfm::View::getID() const { return m_id; }

namespace {

/* Track table as the C ABI exchanges it: offsets[num_tracks + 1]; per feature ids[3] =
 * (view, local id, global id), xy[2], rgb[3]. */
std::vector<orthosfm::Track> build(int64_t num_tracks, const int64_t* offsets, const uint32_t* ids, const float* xy,
                                   const uint32_t* rgb)
{
    std::vector<orthosfm::Track> tracks(static_cast<size_t>(num_tracks));
    for (int64_t t = 0; t < num_tracks; ++t)
        for (int64_t k = offsets[t]; k < offsets[t + 1]; ++k) {
            orthosfm::Feature f(ids[3 * k], ids[3 * k + 1], ids[3 * k + 2], xy[2 * k], xy[2 * k + 1]);
            f.r = rgb[3 * k]; f.g = rgb[3 * k + 1]; f.b = rgb[3 * k + 2];
            tracks[t].add(f);
        }
    return tracks;
}

}  // namespace

extern "C" {

int osfm_refio_save_tracks(const char* path, int64_t num_tracks, const int64_t* offsets, const uint32_t* ids,
                           const float* xy, const uint32_t* rgb)
{
    orthosfm::saveTracksToFile(build(num_tracks, offsets, ids, xy, rgb), path);
    return 0;
}

int osfm_refio_save_pairwise(const char* folder, int num_views, int64_t num_tracks, const int64_t* offsets,
                             const uint32_t* ids, const float* xy, const uint32_t* rgb)
{
    std::vector<orthosfm::View> views;
    for (int v = 0; v < num_views; ++v) views.emplace_back(static_cast<unsigned int>(v), std::string());
    orthosfm::saveTracksToPairwiseFiles(build(num_tracks, offsets, ids, xy, rgb), views, folder);
    return 0;
}

/* Two calls: with out pointers NULL it returns the sizes; then it fills the arrays. */
int osfm_refio_load_tracks(const char* path, int64_t* num_tracks, int64_t* num_features, int64_t* offsets,
                           uint32_t* ids, float* xy, uint32_t* rgb)
{
    std::vector<orthosfm::Track> tracks;
    orthosfm::loadTracksFromFile(tracks, path);
    int64_t n = 0;
    for (size_t t = 0; t < tracks.size(); ++t) {
        if (offsets) offsets[t] = n;
        for (unsigned k = 0; k < tracks[t].size(); ++k, ++n) {
            orthosfm::Feature const& f = tracks[t].get(static_cast<int>(k));
            if (ids) { ids[3 * n] = f.viewID; ids[3 * n + 1] = f.localFeatureID; ids[3 * n + 2] = f.globalFeatureID; }
            if (xy) { xy[2 * n] = f.x; xy[2 * n + 1] = f.y; }
            if (rgb) { rgb[3 * n] = f.r; rgb[3 * n + 1] = f.g; rgb[3 * n + 2] = f.b; }
        }
    }
    if (offsets) offsets[tracks.size()] = n;
    if (num_tracks) *num_tracks = static_cast<int64_t>(tracks.size());
    if (num_features) *num_features = n;
    return 0;
}

}  // extern "C"
