// io_formats.cuh -- the on-disk format downstream MVE tools read the matching result from
// (SURVEY.md section 8, row f4).  Host code only.
//
// What it replaces (reference, paths relative to /root/reference):
//   sfm::bundler::save_prebundle_data / load_prebundle_data / *_to_file / *_from_file
//   (src/mve/sfm/bundler_common.cc:56-190): "MVE_PREBUNDLE\n", then little-endian int32 /
//   float32 / uint8 records: per viewport the feature positions (2 floats each) and colors
//   (3 bytes each), then per matching pair the two view ids and its (i, j) index pairs.
//   orthosfm::saveTracksToFile / loadTracksFromFile / saveTracksToPairwiseFiles
//   (src/matching/matching_io.cpp:16-140): tracks.txt, one line per track,
//   "count;{viewID;localID;globalID;x;y;r;g;b}*" with the iostream default float format
//   (6 significant digits, the printf %g form), and the AAA_BBB.txt files "x1 y1 x2 y2" of
//   the tracks that see both views.  The Feature fields are the ones the MVE bridge fills
//   (src/matching/matching_mve.cpp:455-466): globalID = 32768 * view + feature,
//   x = width * (pos_x + 0.5), y = width * (pos_y + 0.5) computed in double, stored as float.
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace osfm {

constexpr char kPrebundleSignature[] = "MVE_PREBUNDLE\n";    // bundler_common.cc:22-23
constexpr size_t kPrebundleSignatureLen = 14;

struct FileCloser {
    FILE* f;
    ~FileCloser() { if (f) fclose(f); }
};

inline bool write_i32(FILE* f, int32_t v) { return fwrite(&v, sizeof v, 1, f) == 1; }
inline bool read_i32(FILE* f, int32_t* v) { return fread(v, sizeof *v, 1, f) == 1; }

// Returns 0 on success, 1 on an I/O error.
inline int save_prebundle(const char* path, int num_views, const int32_t* features_per_view,
                          const float* positions, const uint8_t* colors, int npairs,
                          const int32_t* pair_views, const int64_t* list_offset, const int32_t* match_ij)
{
    FileCloser fc{fopen(path, "wb")};
    FILE* f = fc.f;
    if (!f) return 1;
    if (fwrite(kPrebundleSignature, 1, kPrebundleSignatureLen, f) != kPrebundleSignatureLen) return 1;
    if (!write_i32(f, num_views)) return 1;
    int64_t at = 0;
    for (int v = 0; v < num_views; ++v) {
        int32_t const n = features_per_view[v];
        // positions and colors are separate vectors in the reference; a matcher-only caller
        // may have neither (counts of 0 are what save_prebundle_data writes for empty vectors)
        int32_t const npos = positions ? n : 0, ncol = colors ? n : 0;
        if (!write_i32(f, npos)) return 1;
        if (npos > 0 && fwrite(positions + 2 * at, sizeof(float) * 2, npos, f) != static_cast<size_t>(npos)) return 1;
        if (!write_i32(f, ncol)) return 1;
        if (ncol > 0 && fwrite(colors + 3 * at, 3, ncol, f) != static_cast<size_t>(ncol)) return 1;
        at += n;
    }
    if (!write_i32(f, npairs)) return 1;
    for (int p = 0; p < npairs; ++p) {
        int64_t const k = list_offset[p + 1] - list_offset[p];
        if (!write_i32(f, pair_views[2 * p]) || !write_i32(f, pair_views[2 * p + 1]) ||
            !write_i32(f, static_cast<int32_t>(k))) return 1;
        if (k > 0 && fwrite(match_ij + 2 * list_offset[p], sizeof(int32_t) * 2, k, f) != static_cast<size_t>(k)) return 1;
    }
    return fflush(f) == 0 ? 0 : 1;
}

struct PrebundleData {
    std::vector<int32_t> n_positions, n_colors;    // per view
    std::vector<float> positions;                  // concatenated, 2 per feature
    std::vector<uint8_t> colors;                   // concatenated, 3 per feature
    std::vector<int32_t> pair_views;               // 2 per pair
    std::vector<int64_t> list_offset;              // npairs + 1
    std::vector<int32_t> match_ij;                 // 2 per match
};

// Returns 0 on success, 1 on an I/O error, 2 on a bad signature / malformed file.
inline int load_prebundle(const char* path, PrebundleData* d)
{
    FileCloser fc{fopen(path, "rb")};
    FILE* f = fc.f;
    if (!f) return 1;
    char sig[kPrebundleSignatureLen];
    if (fread(sig, 1, kPrebundleSignatureLen, f) != kPrebundleSignatureLen) return 2;
    if (memcmp(sig, kPrebundleSignature, kPrebundleSignatureLen) != 0) return 2;
    int32_t nv = 0;
    if (!read_i32(f, &nv) || nv < 0) return 2;
    *d = PrebundleData();
    for (int v = 0; v < nv; ++v) {
        int32_t np = 0, nc = 0;
        if (!read_i32(f, &np) || np < 0) return 2;
        size_t const p0 = d->positions.size();
        d->positions.resize(p0 + 2 * static_cast<size_t>(np));
        if (np > 0 && fread(d->positions.data() + p0, sizeof(float) * 2, np, f) != static_cast<size_t>(np)) return 2;
        if (!read_i32(f, &nc) || nc < 0) return 2;
        size_t const c0 = d->colors.size();
        d->colors.resize(c0 + 3 * static_cast<size_t>(nc));
        if (nc > 0 && fread(d->colors.data() + c0, 3, nc, f) != static_cast<size_t>(nc)) return 2;
        d->n_positions.push_back(np);
        d->n_colors.push_back(nc);
    }
    int32_t npairs = 0;
    if (!read_i32(f, &npairs) || npairs < 0) return 2;
    d->list_offset.push_back(0);
    for (int p = 0; p < npairs; ++p) {
        int32_t a = 0, b = 0, k = 0;
        if (!read_i32(f, &a) || !read_i32(f, &b) || !read_i32(f, &k) || k < 0) return 2;
        d->pair_views.push_back(a);
        d->pair_views.push_back(b);
        size_t const m0 = d->match_ij.size();
        d->match_ij.resize(m0 + 2 * static_cast<size_t>(k));
        if (k > 0 && fread(d->match_ij.data() + m0, sizeof(int32_t) * 2, k, f) != static_cast<size_t>(k)) return 2;
        d->list_offset.push_back(d->list_offset.back() + k);
    }
    return 0;
}

// ---- tracks.txt ------------------------------------------------------------------------------------

struct TrackFeature {
    uint32_t view, local_id, global_id;
    float x, y;
    uint32_t r, g, b;
};

struct TrackTable {
    std::vector<int64_t> offset;            // num_tracks + 1
    std::vector<TrackFeature> features;     // grouped by track
};

// Groups the features by the track id osfm_tracks_compute gave them (-1 = no track): tracks in
// ascending id, features inside a track in ascending (view, feature).  Returns 2 on an id out
// of range.
inline int build_track_table(int num_views, const int32_t* features_per_view, const int32_t* track_of_feature,
                             int num_tracks, const float* positions, double image_width, const uint8_t* colors,
                             TrackTable* t)
{
    int64_t total = 0;
    for (int v = 0; v < num_views; ++v) total += features_per_view[v];
    t->offset.assign(static_cast<size_t>(num_tracks) + 1, 0);
    for (int64_t i = 0; i < total; ++i) {
        int32_t const id = track_of_feature[i];
        if (id < -1 || id >= num_tracks) return 2;
        if (id >= 0) ++t->offset[id + 1];
    }
    for (int k = 0; k < num_tracks; ++k) t->offset[k + 1] += t->offset[k];
    t->features.resize(static_cast<size_t>(t->offset[num_tracks]));
    std::vector<int64_t> fill(t->offset.begin(), t->offset.end() - 1);
    int64_t at = 0;
    for (int v = 0; v < num_views; ++v) {
        for (int f = 0; f < features_per_view[v]; ++f, ++at) {
            int32_t const id = track_of_feature[at];
            if (id < 0) continue;
            TrackFeature& o = t->features[static_cast<size_t>(fill[id]++)];
            o.view = static_cast<uint32_t>(v);
            o.local_id = static_cast<uint32_t>(f);
            o.global_id = 32768u * static_cast<uint32_t>(v) + static_cast<uint32_t>(f);
            o.x = positions ? static_cast<float>(image_width * (positions[2 * at] + 0.5)) : 0.0f;
            o.y = positions ? static_cast<float>(image_width * (positions[2 * at + 1] + 0.5)) : 0.0f;
            o.r = colors ? colors[3 * at] : 0;
            o.g = colors ? colors[3 * at + 1] : 0;
            o.b = colors ? colors[3 * at + 2] : 0;
        }
    }
    return 0;
}

inline int save_tracks(const char* path, const TrackTable& t)
{
    FileCloser fc{fopen(path, "w")};
    FILE* f = fc.f;
    if (!f) return 1;
    size_t const ntracks = t.offset.size() - 1;
    for (size_t k = 0; k < ntracks; ++k) {
        int64_t const b = t.offset[k], e = t.offset[k + 1];
        if (fprintf(f, "%lld;", static_cast<long long>(e - b)) < 0) return 1;
        for (int64_t i = b; i < e; ++i) {
            TrackFeature const& o = t.features[static_cast<size_t>(i)];
            if (fprintf(f, "%u;%u;%u;%g;%g;%u;%u;%u%s", o.view, o.local_id, o.global_id, o.x, o.y, o.r, o.g, o.b,
                        i + 1 < e ? ";" : "") < 0) return 1;
        }
        if (fputc('\n', f) == EOF) return 1;
    }
    return fflush(f) == 0 ? 0 : 1;
}

// Returns 0, 1 (cannot open) or 2 (a line that does not hold count * 8 fields).
inline int load_tracks(const char* path, TrackTable* t)
{
    FileCloser fc{fopen(path, "r")};
    FILE* f = fc.f;
    if (!f) return 1;
    t->offset.assign(1, 0);
    t->features.clear();
    std::string line;
    std::vector<char> buf(1 << 16);
    while (true) {
        line.clear();
        bool got = false;
        while (fgets(buf.data(), static_cast<int>(buf.size()), f)) {
            got = true;
            line += buf.data();
            if (!line.empty() && line.back() == '\n') break;
        }
        if (!got) break;
        while (!line.empty() && (line.back() == '\n' || line.back() == '\r')) line.pop_back();
        const char* p = line.c_str();
        char* end = nullptr;
        long const count = strtol(p, &end, 10);
        if (end == p || count < 0) return 2;
        p = end;
        for (long k = 0; k < count; ++k) {
            TrackFeature o{};
            double vals[8];
            for (int q = 0; q < 8; ++q) {
                if (*p != ';') return 2;
                ++p;
                vals[q] = strtod(p, &end);
                if (end == p) return 2;
                p = end;
            }
            o.view = static_cast<uint32_t>(vals[0]);
            o.local_id = static_cast<uint32_t>(vals[1]);
            o.global_id = static_cast<uint32_t>(vals[2]);
            o.x = static_cast<float>(vals[3]);
            o.y = static_cast<float>(vals[4]);
            o.r = static_cast<uint32_t>(vals[5]);
            o.g = static_cast<uint32_t>(vals[6]);
            o.b = static_cast<uint32_t>(vals[7]);
            t->features.push_back(o);
        }
        t->offset.push_back(static_cast<int64_t>(t->features.size()));
    }
    return 0;
}

// AAA_BBB.txt for every pair of views (view_ids ascending as given) that shares a track: one
// line "x1 y1 x2 y2" per track that holds exactly one feature of each of the two views
// (filterTracksToAvailableCameras(ids, tracks, true, false), src/util/common.cpp:85-130).
// Returns the number of files written, or -1 on an I/O error.
inline int save_pairwise_tracks(const char* folder, const TrackTable& t, int num_views, const int32_t* view_ids)
{
    size_t const ntracks = t.offset.size() - 1;
    int written = 0;
    for (int a = 0; a < num_views; ++a) {
        for (int b = a + 1; b < num_views; ++b) {
            uint32_t const ia = static_cast<uint32_t>(view_ids[a]), ib = static_cast<uint32_t>(view_ids[b]);
            FILE* f = nullptr;
            for (size_t k = 0; k < ntracks; ++k) {
                int na = 0, nb = 0;
                for (int64_t i = t.offset[k]; i < t.offset[k + 1]; ++i) {
                    na += t.features[static_cast<size_t>(i)].view == ia;
                    nb += t.features[static_cast<size_t>(i)].view == ib;
                }
                if (na + nb != 2) continue;       // "full size": exactly ids.size() features survive the filter
                if (!f) {
                    char name[64];
                    snprintf(name, sizeof name, "/%03d_%03d.txt", view_ids[a], view_ids[b]);
                    f = fopen((std::string(folder) + name).c_str(), "w");
                    if (!f) return -1;
                    ++written;
                }
                // first the features seen by view a, then those seen by view b, as the reference loops
                for (int side = 0; side < 2; ++side) {
                    uint32_t const id = side == 0 ? ia : ib;
                    for (int64_t i = t.offset[k]; i < t.offset[k + 1]; ++i) {
                        TrackFeature const& o = t.features[static_cast<size_t>(i)];
                        if (o.view != id) continue;
                        fprintf(f, side == 0 ? "%g %g " : "%g %g\n", o.x, o.y);
                    }
                }
            }
            if (f && fclose(f) != 0) return -1;
        }
    }
    return written;
}

}  // namespace osfm
