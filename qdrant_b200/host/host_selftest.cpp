// host_selftest.cpp — drives libqdrant_b200.so through the C++ mirror of the reference's interface
// (RawScorerBuilder -> RawScorer, FilteredScorer, BatchFilteredSearcher).  Inputs/outputs are raw binary files so that
// tests/test_gpu_cpp_host.py can compare the results with the CPU restatement bit for bit.
//   host_selftest <base.f32> <queries.f32> <n> <dim> <nq> <top> <out.bin>
// out.bin: for each query: u32 count, count x {u32 idx, f32 score}; then nq x 32 f32 = score_points of ids 0..31 after
// FilteredScorer filtering (every 3rd point deleted), then 1 f32 = score_internal(1, 2).
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>

#include "qdrant_b200.hpp"

using namespace qdrant_b200;

static std::vector<float> read_f32(const char* path, size_t n) {
    std::vector<float> v(n);
    std::ifstream f(path, std::ios::binary);
    f.read(reinterpret_cast<char*>(v.data()), (std::streamsize)(n * 4));
    if (!f) { std::cerr << "short read: " << path << "\n"; std::exit(2); }
    return v;
}

int main(int argc, char** argv) {
    if (argc != 8) { std::cerr << "usage: host_selftest base queries n dim nq top out\n"; return 2; }
    const size_t n = std::strtoull(argv[3], nullptr, 10), dim = std::strtoull(argv[4], nullptr, 10);
    const uint32_t nq = (uint32_t)std::strtoul(argv[5], nullptr, 10), top = (uint32_t)std::strtoul(argv[6], nullptr, 10);
    try {
        auto base = read_f32(argv[1], n * dim);
        auto queries = read_f32(argv[2], (size_t)nq * dim);
        auto storage = VectorStorage::dense_f32(0, Distance::Dot, (uint32_t)dim, n, base.data());
        std::ofstream out(argv[7], std::ios::binary);
        BatchFilteredSearcher searcher(queries.data(), nq, *storage, top);
        auto res = searcher.peek_top_all();
        for (auto& r : res) {
            uint32_t c = (uint32_t)r.size();
            out.write(reinterpret_cast<const char*>(&c), 4);
            out.write(reinterpret_cast<const char*>(r.data()), (std::streamsize)(r.size() * sizeof(ScoredPointOffset)));
        }
        std::vector<bool> deleted(n, false);
        for (size_t i = 0; i < n; i += 3) deleted[i] = true;
        for (uint32_t q = 0; q < nq; ++q) {
            FilteredScorer fs(storage->build_raw_scorer(queries.data() + (size_t)q * dim), &deleted);
            std::vector<PointOffsetType> ids;
            for (PointOffsetType i = 0; i < 48; ++i) ids.push_back(i);
            auto scored = fs.score_points(ids, 32);
            std::vector<float> s(32, 0.f);
            for (size_t i = 0; i < scored.size() && i < 32; ++i) s[i] = scored[i].score;
            out.write(reinterpret_cast<const char*>(s.data()), 32 * 4);
        }
        auto sc = storage->build_raw_scorer(queries.data());
        float si = sc->score_internal(1, 2);
        out.write(reinterpret_cast<const char*>(&si), 4);
        // error behaviour: out-of-range id is a construction/argument error, never a silent result
        bool threw = false;
        try { sc->score_point((PointOffsetType)n + 5); } catch (const OperationError& e) { threw = e.status == QB_ERR_INVALID; }
        if (!threw) { std::cerr << "expected QB_ERR_INVALID for out-of-range id\n"; return 3; }
    } catch (const std::exception& e) {
        std::cerr << "error: " << e.what() << "\n";
        return 1;
    }
    return 0;
}
