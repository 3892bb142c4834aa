// qdrant_b200.hpp — C++ host side above the C ABI (include/qb200.h), mirroring the reference's scorer interface.
//
// The reference is compiled (Rust) code and its toolchain is absent from the build image, so the host-side mirror of
// its operator interface is written in C++ (header only; links against libqdrant_b200.so):
//     RawScorer               lib/segment/src/vector_storage/raw_scorer.rs:39-54
//     RawScorerBuilder        raw_scorer.rs:122-128
//     FilteredScorer          lib/segment/src/index/hnsw_index/point_scorer.rs:53-58,160-304
//     BatchFilteredSearcher   point_scorer.rs:312-472
//     ScoredPointOffset       lib/common/common/src/types.rs:12-31
//     GraphLayers::search     lib/segment/src/index/hnsw_index/graph_layers.rs:530-561   (HnswGraph: whole batches on the device)
//     SegmentsSearcher task + BatchResultAggregator   segments_searcher.rs:255, search_result_aggregator.rs:50-117   (SegmentShard)
// Same names, argument meaning and error behaviour: construction may fail (OperationError -> std::runtime_error),
// scoring is infallible apart from device failures (which throw, the analogue of the reference's `expect`).
#pragma once
#include <atomic>
#include <cstdint>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/qb200.h"

namespace qdrant_b200 {

using PointOffsetType = uint32_t;  // common::types::PointOffsetType
using ScoreType = float;           // common::types::ScoreType
using ScoredPointOffset = qb_scored_point;
constexpr size_t VECTOR_READ_BATCH_SIZE = 64;  // vector_storage/common.rs:20

struct OperationError : std::runtime_error {
    qb_status status;
    OperationError(qb_status s, const std::string& what) : std::runtime_error(what), status(s) {}
};
inline void check(qb_status st) {
    if (st != QB_OK) throw OperationError(st, qb_last_error());
}

enum class Distance : int { Cosine = QB_DIST_COSINE, Euclid = QB_DIST_EUCLID, Dot = QB_DIST_DOT, Manhattan = QB_DIST_MANHATTAN };

// Box<dyn RawScorer + 'a>
class RawScorer {
public:
    virtual ~RawScorer() = default;
    virtual void score_points(const PointOffsetType* points, size_t n, ScoreType* scores) const = 0;
    virtual ScoreType score_point(PointOffsetType point) const = 0;
    virtual ScoreType score_internal(PointOffsetType a, PointOffsetType b) const = 0;
};

class VectorStorage;

class B200RawScorer final : public RawScorer {
public:
    B200RawScorer(qb_scorer* h) : h_(h) {}
    ~B200RawScorer() override { qb_scorer_destroy(h_); }
    B200RawScorer(const B200RawScorer&) = delete;
    void score_points(const PointOffsetType* points, size_t n, ScoreType* scores) const override { check(qb_score_points(h_, points, n, scores)); }
    ScoreType score_point(PointOffsetType point) const override {
        ScoreType s;
        check(qb_score_point(h_, point, &s));
        return s;
    }
    ScoreType score_internal(PointOffsetType a, PointOffsetType b) const override {
        ScoreType s;
        check(qb_score_internal(h_, a, b, &s));
        return s;
    }
    qb_scorer* raw() const { return h_; }

private:
    qb_scorer* h_;
};

// A segment's vectors resident in HBM; doubles as RawScorerBuilder / QuantizedVectorsRead.
class VectorStorage {
public:
    explicit VectorStorage(qb_storage* h) : h_(h) {}
    ~VectorStorage() { qb_storage_destroy(h_); }
    VectorStorage(const VectorStorage&) = delete;

    static std::unique_ptr<VectorStorage> dense_f32(int device, Distance d, uint32_t dim, uint64_t count, const float* rows) {
        qb_storage* h = nullptr;
        check(qb_storage_create_dense(device, QB_DT_F32, static_cast<qb_distance>(d), dim, count, rows, (uint64_t)dim * 4, &h));
        return std::make_unique<VectorStorage>(h);
    }
    static std::unique_ptr<VectorStorage> sq8(int device, Distance d, uint32_t dim, uint64_t count, const uint8_t* rows, uint32_t row_bytes, float alpha,
                                             float offset, float multiplier) {
        // construct_vector_parameters (quantized_vectors.rs:205-234)
        const qb_qdistance dt = (d == Distance::Euclid) ? QB_QD_L2 : (d == Distance::Manhattan ? QB_QD_L1 : QB_QD_DOT);
        const int invert = (d == Distance::Euclid || d == Distance::Manhattan) ? 1 : 0;
        qb_storage* h = nullptr;
        check(qb_storage_create_sq8(device, dim, count, rows, row_bytes, alpha, offset, multiplier, dt, invert, static_cast<qb_distance>(d), &h));
        return std::make_unique<VectorStorage>(h);
    }
    // RawScorerBuilder::build_raw_scorer / QuantizedVectorsRead::raw_scorer
    std::unique_ptr<RawScorer> build_raw_scorer(const float* query) const {
        qb_scorer* sc = nullptr;
        check(qb_scorer_create(h_, query, &sc));
        return std::make_unique<B200RawScorer>(sc);
    }
    // new_raw_scorer for QueryVector::{RecommendBestScore, RecommendSumScores, Discover, Context} (raw_scorer.rs:228-333):
    // `vectors` = the flattened example vectors in the layout include/qb200.h documents
    std::unique_ptr<RawScorer> build_custom_scorer(qb_query_kind kind, const float* vectors, uint32_t n_a, uint32_t n_b) const {
        qb_scorer* sc = nullptr;
        check(qb_scorer_create_custom(h_, kind, vectors, n_a, n_b, &sc));
        return std::make_unique<B200RawScorer>(sc);
    }
    // QuantizedVectorsRead::raw_internal_scorer (throws OperationError{QB_ERR_UNSUPPORTED} for PQ)
    std::unique_ptr<RawScorer> raw_internal_scorer(PointOffsetType point) const {
        qb_scorer* sc = nullptr;
        check(qb_scorer_create_internal(h_, point, &sc));
        return std::make_unique<B200RawScorer>(sc);
    }
    uint64_t count() const {
        uint64_t c = 0;
        check(qb_storage_info(h_, nullptr, &c, nullptr));
        return c;
    }
    uint32_t dim() const {
        uint32_t d = 0;
        check(qb_storage_info(h_, &d, nullptr, nullptr));
        return d;
    }
    qb_storage* raw() const { return h_; }

private:
    qb_storage* h_;
};

// point_scorer.rs:53-58: RawScorer + filters (deleted bitslice, optional payload filter)
class FilteredScorer {
public:
    FilteredScorer(std::unique_ptr<RawScorer> raw, const std::vector<bool>* point_deleted = nullptr,
                   std::function<bool(PointOffsetType)> filter = nullptr)
        : raw_(std::move(raw)), deleted_(point_deleted), filter_(std::move(filter)) {}

    bool check_vector(PointOffsetType p) const {
        if (deleted_ && p < deleted_->size() && (*deleted_)[p]) return false;
        return !filter_ || filter_(p);
    }
    // point_scorer.rs:265-295: filters `point_ids` in place, truncates to `limit` (0 = no limit), scores in ONE batch call
    std::vector<ScoredPointOffset> score_points(std::vector<PointOffsetType>& point_ids, size_t limit) {
        size_t kept = 0;
        for (PointOffsetType p : point_ids)
            if (check_vector(p)) point_ids[kept++] = p;
        point_ids.resize(kept);
        if (limit != 0 && point_ids.size() > limit) point_ids.resize(limit);
        scores_.resize(point_ids.size());
        raw_->score_points(point_ids.data(), point_ids.size(), scores_.data());
        std::vector<ScoredPointOffset> out(point_ids.size());
        for (size_t i = 0; i < point_ids.size(); ++i) out[i] = ScoredPointOffset{point_ids[i], scores_[i]};
        return out;
    }
    ScoreType score_point(PointOffsetType p) const { return raw_->score_point(p); }
    ScoreType score_internal(PointOffsetType a, PointOffsetType b) const { return raw_->score_internal(a, b); }

private:
    std::unique_ptr<RawScorer> raw_;
    const std::vector<bool>* deleted_;
    std::function<bool(PointOffsetType)> filter_;
    std::vector<ScoreType> scores_;
};

// point_scorer.rs:312-472: the reference walks 64-id chunks x per-query scorers x per-query heaps; here the scan and the
// top-k are one fused library call.
class BatchFilteredSearcher {
public:
    BatchFilteredSearcher(const float* queries, uint32_t n_queries, const VectorStorage& storage, uint32_t top, const uint64_t* point_deleted_words = nullptr)
        : queries_(queries), nq_(n_queries), st_(storage), top_(top), deleted_(point_deleted_words) {
        if (top == 0) throw std::invalid_argument("length must be greater than zero");  // FixedLengthPriorityQueue::new
    }
    // peek_top_all: every non-deleted point; is_stopped mirrors &AtomicBool (check_process_stopped)
    std::vector<std::vector<ScoredPointOffset>> peek_top_all(const std::atomic<int32_t>* is_stopped = nullptr) const { return run(nullptr, 0, is_stopped); }
    // peek_top_iter over an explicit id list (deferred / filtered points already removed by the caller)
    std::vector<std::vector<ScoredPointOffset>> peek_top_iter(const std::vector<PointOffsetType>& points, const std::atomic<int32_t>* is_stopped = nullptr) const {
        return run(points.data(), points.size(), is_stopped);
    }

private:
    std::vector<std::vector<ScoredPointOffset>> run(const PointOffsetType* ids, uint64_t n_ids, const std::atomic<int32_t>* is_stopped) const {
        std::vector<ScoredPointOffset> flat((size_t)nq_ * top_);
        std::vector<uint32_t> counts(nq_);
        static const PointOffsetType kNone = 0;
        check(qb_search_batch(st_.raw(), queries_, nq_, top_, deleted_, (ids || n_ids) ? (ids ? ids : &kNone) : nullptr, n_ids,
                              reinterpret_cast<const volatile int32_t*>(is_stopped), flat.data(), counts.data(), nullptr));
        std::vector<std::vector<ScoredPointOffset>> out(nq_);
        for (uint32_t q = 0; q < nq_; ++q) out[q].assign(flat.begin() + (size_t)q * top_, flat.begin() + (size_t)q * top_ + counts[q]);
        return out;
    }
    const float* queries_;
    uint32_t nq_;
    const VectorStorage& st_;
    uint32_t top_;
    const uint64_t* deleted_;
};

// GraphLayers::search (index/hnsw_index/graph_layers.rs:530-561) for a batch of queries, traversal and scoring on the device.
// The graph is the segment's links.bin in GraphLinksFormat::Plain (graph_links/view.rs:121-135).
class HnswGraph {
public:
    HnswGraph(const VectorStorage& storage, const uint8_t* links_bin, uint64_t n_bytes, uint32_t m, uint32_t m0) {
        check(qb_hnsw_create_plain(storage.raw(), links_bin, n_bytes, m, m0, &h_));
    }
    ~HnswGraph() { qb_hnsw_destroy(h_); }
    HnswGraph(const HnswGraph&) = delete;
    // entry_point / entry_level = GraphLayers::get_entry_point(filters, custom_entry_points); deleted = the filter as a bitmap (bit = 1: skip)
    std::vector<std::vector<ScoredPointOffset>> search(const float* queries, uint32_t n_queries, uint32_t top, uint32_t ef, PointOffsetType entry_point,
                                                       uint32_t entry_level, const uint64_t* deleted = nullptr) const {
        std::vector<ScoredPointOffset> flat((size_t)n_queries * top);
        std::vector<uint32_t> counts(n_queries);
        check(qb_hnsw_search_batch(h_, queries, n_queries, top, ef, entry_point, entry_level, deleted, nullptr, flat.data(), counts.data(), nullptr));
        std::vector<std::vector<ScoredPointOffset>> out(n_queries);
        for (uint32_t q = 0; q < n_queries; ++q) out[q].assign(flat.begin() + (size_t)q * top, flat.begin() + (size_t)q * top + counts[q]);
        return out;
    }

private:
    qb_hnsw* h_ = nullptr;
};

// One shard (segment on one GPU) of a sharded search: SegmentsSearcher's blocking task per segment (segments_searcher.rs:255) calls
// search() with the same queries on every shard; the BatchResultAggregator step (search_result_aggregator.rs:50-117) happens on the
// devices and every call returns the merged top-k.  Wire the shards of one process with ShardedSegments::connect.
class SegmentShard {
public:
    SegmentShard(const VectorStorage& storage, int device, int rank, int world, uint32_t max_queries, uint32_t max_top) : st_(storage) {
        check(qb_comm_create(device, rank, world, max_queries, max_top, &c_));
    }
    ~SegmentShard() { qb_comm_destroy(c_); }
    SegmentShard(const SegmentShard&) = delete;
    static void connect(const std::vector<SegmentShard*>& shards) {
        std::vector<qb_comm*> cs;
        for (auto* s : shards) cs.push_back(s->c_);
        check(qb_comm_connect_local(cs.data(), (int32_t)cs.size()));
    }
    std::vector<std::vector<ScoredPointOffset>> search(const float* queries, uint32_t n_queries, uint32_t top) const {
        std::vector<ScoredPointOffset> flat((size_t)n_queries * top);
        std::vector<uint32_t> counts(n_queries);
        check(qb_multi_search_batch(c_, st_.raw(), queries, n_queries, top, nullptr, nullptr, flat.data(), counts.data(), nullptr));
        std::vector<std::vector<ScoredPointOffset>> out(n_queries);
        for (uint32_t q = 0; q < n_queries; ++q) out[q].assign(flat.begin() + (size_t)q * top, flat.begin() + (size_t)q * top + counts[q]);
        return out;
    }

private:
    const VectorStorage& st_;
    qb_comm* c_ = nullptr;
};

}  // namespace qdrant_b200
