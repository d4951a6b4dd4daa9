// Stand-in for boost::split / boost::is_any_of (see README.md), as loadTracksFromFile uses them:
// split at any of the given characters, empty tokens kept (token_compress_off).
#pragma once
#include <string>
#include <vector>
namespace boost {
struct any_of_pred { std::string set; bool operator()(char c) const { return set.find(c) != std::string::npos; } };
inline any_of_pred is_any_of(const char* s) { return any_of_pred{s}; }
template <typename Pred>
inline void split(std::vector<std::string>& out, const std::string& in, Pred pred) {
    out.clear();
    std::string cur;
    for (char c : in) {
        if (pred(c)) { out.push_back(cur); cur.clear(); } else cur.push_back(c);
    }
    out.push_back(cur);
}
}  // namespace boost
