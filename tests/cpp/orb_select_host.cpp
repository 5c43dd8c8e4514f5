// Host build of epivo_b200/csrc/orb_select.cuh next to the real thing: OpenCV's KeyPointsFilter::retainBest spelled with
// std::nth_element / std::partition on KeyPoint-like records (keypoint.cpp), and libstdc++'s own std::__heap_select.
// tests/test_orb_select.py compares the two (and the numpy oracle) on random and tie-heavy inputs.
#include <algorithm>
#include <vector>

#include "../../epivo_b200/csrc/orb_select.cuh"

namespace {
struct Kp {
    float response;
    int idx;
};
struct ResponseGreater {
    bool operator()(const Kp& a, const Kp& b) const { return a.response > b.response; }
};
struct ResponseGreaterEq {
    float value;
    bool operator()(const Kp& k) const { return k.response >= value; }
};
}  // namespace

extern "C" {

// KeyPointsFilter::retainBest as OpenCV writes it; out_idx receives the kept original indices in order
int ref_retain_best(const uint8_t* resp, int n, int n_points, int* out_idx) {
    std::vector<Kp> k(n);
    for (int i = 0; i < n; ++i) k[i] = Kp{(float)resp[i], i};
    if (n_points >= 0 && k.size() > (size_t)n_points) {
        if (n_points == 0) {
            k.clear();
        } else {
            std::nth_element(k.begin(), k.begin() + n_points - 1, k.end(), ResponseGreater());
            const float amb = k[n_points - 1].response;
            auto new_end = std::partition(k.begin() + n_points, k.end(), ResponseGreaterEq{amb});
            k.resize(new_end - k.begin());
        }
    }
    for (size_t i = 0; i < k.size(); ++i) out_idx[i] = k[i].idx;
    return (int)k.size();
}

int epv_retain_best_host(const uint8_t* resp, int n, int n_points, int* out_idx) {
    std::vector<uint32_t> k(n);
    for (int i = 0; i < n; ++i) k[i] = ((uint32_t)resp[i] << 24) | (uint32_t)i;
    int kept = n;
    if (n > n_points) kept = n_points == 0 ? 0 : orb_retain_best(k.data(), n, n_points);
    for (int i = 0; i < kept; ++i) out_idx[i] = (int)(k[i] & 0xFFFFFFu);
    return kept;
}

// the block-parallel form of the same selection (what orb_select_kernel runs), executed by one host "thread"; depth < 0
// keeps libstdc++'s 2 lg n, depth >= 0 forces the heap-select fallback after that many partition passes
int epv_retain_best_block_host(const uint8_t* resp, int n, int n_points, int* out_idx) {
    std::vector<uint32_t> k(n), A(n + 1), B(n + 1);
    for (int i = 0; i < n; ++i) k[i] = ((uint32_t)resp[i] << 24) | (uint32_t)i;
    int kept = n;
    OrbSelectHostExec ex;
    OrbSelectShared sh;
    if (n > n_points) kept = n_points == 0 ? 0 : orb_retain_best_block(ex, k.data(), n, n_points, A.data(), B.data(), &sh);
    for (int i = 0; i < kept; ++i) out_idx[i] = (int)(k[i] & 0xFFFFFFu);
    return kept;
}

// the depth-limit fallback of std::nth_element, which random inputs never reach: both forms permute all n entries
void ref_heap_select(const uint8_t* resp, int n, int first, int middle, int* out_idx) {
    std::vector<Kp> k(n);
    for (int i = 0; i < n; ++i) k[i] = Kp{(float)resp[i], i};
    std::__heap_select(k.begin() + first, k.begin() + middle, k.end(), __gnu_cxx::__ops::__iter_comp_iter(ResponseGreater()));
    for (int i = 0; i < n; ++i) out_idx[i] = k[i].idx;
}

void epv_heap_select_host(const uint8_t* resp, int n, int first, int middle, int* out_idx) {
    std::vector<uint32_t> k(n);
    for (int i = 0; i < n; ++i) k[i] = ((uint32_t)resp[i] << 24) | (uint32_t)i;
    orb_heap_select(k.data(), first, middle, n);
    for (int i = 0; i < n; ++i) out_idx[i] = (int)(k[i] & 0xFFFFFFu);
}
}
