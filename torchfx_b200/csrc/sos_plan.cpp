// sos_plan.cpp -- see sos_plan.h.
//
// Why a plan exists.  The reference parallelises one channel over time with a Blelloch
// scan of 2x3 matrices through global memory (cuda/parallel_scan.cu:117-215, ~40 B/sample
// /section of f64 traffic).  This engine instead cuts every channel into S independent
// time segments ("streams"), each filtered sequentially by one thread with the whole
// cascade state in registers.  A segment that does not start at n = 0 obtains its initial
// state by first running the same recurrence over the `warm` samples that precede it,
// starting from silence: the state error left after w samples is A^w applied to the true
// state, so once ||A^w|| is below 2^-30 (f32) / 2^-42 / 2^-62 (f64) the segment is
// indistinguishable from an unbroken sequential run at the working precision.  The bound
// is computed here from the actual coefficients (matrix powers by repeated squaring, in
// double); a cascade that does not decay (|pole| >= 1) is simply never split.
#include "sos_plan.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <list>
#include <mutex>
#include <string>
#include <unordered_map>

#include "common.cuh"

namespace tfx {
namespace {

using Mat = std::vector<double>;  // D x D row-major

Mat matmul(const Mat &a, const Mat &b, int D) {
    Mat c(static_cast<size_t>(D) * D, 0.0);
    for (int i = 0; i < D; ++i)
        for (int k = 0; k < D; ++k) {
            const double aik = a[i * D + k];
            if (aik == 0.0) continue;
            for (int j = 0; j < D; ++j) c[i * D + j] += aik * b[k * D + j];
        }
    return c;
}

double norm_inf(const Mat &a, int D) {
    double m = 0.0;
    for (int i = 0; i < D; ++i) {
        double s = 0.0;
        for (int j = 0; j < D; ++j) s += std::fabs(a[i * D + j]);
        if (!(s <= m)) m = s;  // also propagates NaN
    }
    return m;
}

// One homogeneous (x = 0) step of the transposed-direct-form-II cascade, the recurrence
// the kernels run:  y = b0*x + s1;  s1' = b1*x - a1*y + s2;  s2' = b2*x - a2*y;  x_next = y.
void df2t_step(const SosSection *sec, int k, double *s /* 2k */, double x) {
    for (int i = 0; i < k; ++i) {
        const SosSection &c = sec[i];
        const double y = c.b0 * x + s[2 * i];
        s[2 * i] = c.b1 * x - c.a1 * y + s[2 * i + 1];
        s[2 * i + 1] = c.b2 * x - c.a2 * y;
        x = y;
    }
}

constexpr int kMaxLog2 = 26;  // segments never need more than 2^26 warm-up samples; beyond: no split

void analyse_pass(const SosSection *sec, SosPass &p) {
    const int D = 2 * p.k;
    Mat A(static_cast<size_t>(D) * D, 0.0);
    std::vector<double> s(D);
    for (int j = 0; j < D; ++j) {
        std::fill(s.begin(), s.end(), 0.0);
        s[j] = 1.0;
        df2t_step(sec, p.k, s.data(), 0.0);
        for (int i = 0; i < D; ++i) A[i * D + j] = s[i];
    }
    std::vector<Mat> P;
    std::vector<double> nrm;
    P.push_back(A);
    nrm.push_back(norm_inf(A, D));
    double growth = std::max(1.0, nrm[0]);
    for (int i = 1; i <= kMaxLog2; ++i) {
        P.push_back(matmul(P.back(), P.back(), D));
        nrm.push_back(norm_inf(P.back(), D));
        if (!std::isfinite(nrm.back())) break;
        growth = std::max(growth, nrm.back());
        if (nrm.back() < 1e-300) break;
    }
    auto first_below = [&](double tol) -> int64_t {
        if (!std::isfinite(nrm.back())) return -1;
        const double t = tol / growth;
        int hi = -1;
        for (size_t i = 0; i < nrm.size(); ++i)
            if (nrm[i] <= t) {
                hi = static_cast<int>(i);
                break;
            }
        if (hi < 0) return -1;
        // largest n < 2^hi with ||A^n|| > t, greedy over the stored powers
        int64_t n = 0;
        Mat M;
        bool have = false;
        for (int b = hi - 1; b >= 0; --b) {
            Mat Mt = have ? matmul(M, P[b], D) : P[b];
            if (norm_inf(Mt, D) > t) {
                M.swap(Mt);
                have = true;
                n += int64_t(1) << b;
            }
        }
        // ||A^n|| is not monotone in n: keep a margin.
        return n + 1 + n / 8 + 8;
    };
    p.warm_f32 = first_below(std::ldexp(1.0, -30));
    p.warm_f64_io32 = first_below(std::ldexp(1.0, -42));
    p.warm_f64_io64 = first_below(std::ldexp(1.0, -62));
}

// TFX_PREC_AUTO policy: run a short broadband probe through (a) the reference's arithmetic
// (f64 DF1, cpu/iir_cpu.cpp:132-147) and (b) the float32 DF2T recurrence of the fast
// kernel, and keep float32 only if its error stays under kAutoF32Bound of max|y|.
constexpr double kAutoF32Bound = 2.0e-6;
constexpr int kProbeLen = 8192;

void probe_precision(SosPlan &plan) {
    const int K = plan.K;
    std::vector<double> sx0(K, 0.0), sx1(K, 0.0), sy0(K, 0.0), sy1(K, 0.0);
    std::vector<float> t1(K, 0.f), t2(K, 0.f);
    std::vector<float> b0(K), b1(K), b2(K), a1(K), a2(K);
    for (int k = 0; k < K; ++k) {
        b0[k] = static_cast<float>(plan.sec[k].b0);
        b1[k] = static_cast<float>(plan.sec[k].b1);
        b2[k] = static_cast<float>(plan.sec[k].b2);
        a1[k] = static_cast<float>(plan.sec[k].a1);
        a2[k] = static_cast<float>(plan.sec[k].a2);
    }
    uint64_t lcg = 0x9E3779B97F4A7C15ull;
    double max_y = 0.0, max_err = 0.0;
    bool finite = true;
    for (int n = 0; n < kProbeLen; ++n) {
        lcg = lcg * 6364136223846793005ull + 1442695040888963407ull;
        const float xf = static_cast<float>(static_cast<double>(static_cast<int64_t>(lcg >> 11)) *
                                                (1.0 / 4503599627370496.0) -
                                            1.0);  // uniform [-1, 1)
        double v = xf;
        for (int k = 0; k < K; ++k) {
            const SosSection &c = plan.sec[k];
            const double y = c.b0 * v + c.b1 * sx0[k] + c.b2 * sx1[k] - c.a1 * sy0[k] - c.a2 * sy1[k];
            sx1[k] = sx0[k];
            sx0[k] = v;
            sy1[k] = sy0[k];
            sy0[k] = y;
            v = y;
        }
        float u = xf;
        for (int k = 0; k < K; ++k) {
            const float y = std::fmaf(b0[k], u, t1[k]);
            t1[k] = std::fmaf(-a1[k], y, std::fmaf(b1[k], u, t2[k]));
            t2[k] = std::fmaf(-a2[k], y, b2[k] * u);
            u = y;
        }
        if (!std::isfinite(v) || !std::isfinite(u)) {
            finite = false;
            break;
        }
        max_y = std::max(max_y, std::fabs(v));
        max_err = std::max(max_err, std::fabs(static_cast<double>(u) - v));
    }
    if (!finite || max_y == 0.0) {
        plan.probe_rel_err = finite ? 0.0 : INFINITY;
        plan.auto_prec = finite ? TFX_PREC_F32 : TFX_PREC_F64;
        return;
    }
    plan.probe_rel_err = max_err / max_y;
    plan.auto_prec = plan.probe_rel_err <= kAutoF32Bound ? TFX_PREC_F32 : TFX_PREC_F64;
}

// Probe of the mixed-precision recurrence the tile kernel runs (sos_tile.cu, Cascade<MixedF>):
// sections in `mask` in float64 with a float32 signal in and out, the others in float32.
double probe_mixed(const SosPlan &plan, uint64_t mask, int len) {
    const int K = plan.K;
    std::vector<double> sx0(K, 0.0), sx1(K, 0.0), sy0(K, 0.0), sy1(K, 0.0), d1(K, 0.0), d2(K, 0.0);
    std::vector<float> f1(K, 0.f), f2(K, 0.f);
    uint64_t lcg = 0x9E3779B97F4A7C15ull;
    double max_y = 0.0, max_err = 0.0;
    for (int n = 0; n < len; ++n) {
        lcg = lcg * 6364136223846793005ull + 1442695040888963407ull;
        const float xf = static_cast<float>(static_cast<double>(static_cast<int64_t>(lcg >> 11)) * (1.0 / 4503599627370496.0) - 1.0);
        double v = xf;
        for (int k = 0; k < K; ++k) {
            const SosSection &c = plan.sec[k];
            const double y = c.b0 * v + c.b1 * sx0[k] + c.b2 * sx1[k] - c.a1 * sy0[k] - c.a2 * sy1[k];
            sx1[k] = sx0[k];
            sx0[k] = v;
            sy1[k] = sy0[k];
            sy0[k] = y;
            v = y;
        }
        float u = xf;
        for (int k = 0; k < K; ++k) {
            const SosSection &c = plan.sec[k];
            if ((mask >> k) & 1ull) {
                const double ud = u;
                const double y = std::fma(c.b0, ud, d1[k]);
                d1[k] = std::fma(-c.a1, y, std::fma(c.b1, ud, d2[k]));
                d2[k] = std::fma(-c.a2, y, c.b2 * ud);
                u = static_cast<float>(y);
            } else {
                const float b0 = static_cast<float>(c.b0), b1 = static_cast<float>(c.b1), b2 = static_cast<float>(c.b2);
                const float na1 = static_cast<float>(-c.a1), na2 = static_cast<float>(-c.a2);
                const float y = std::fmaf(b0, u, f1[k]);
                f1[k] = std::fmaf(na1, y, std::fmaf(b1, u, f2[k]));
                f2[k] = std::fmaf(na2, y, b2 * u);
                u = y;
            }
        }
        if (!std::isfinite(v) || !std::isfinite(u)) return INFINITY;
        max_y = std::max(max_y, std::fabs(v));
        max_err = std::max(max_err, std::fabs(static_cast<double>(u) - v));
    }
    return max_y > 0.0 ? max_err / max_y : 0.0;
}

// Which sections have to be float64?  Rank the sections by how much promoting each one alone
// helps, then promote in that order until the cascade meets the bound.
void choose_mixed_mask(SosPlan &plan) {
    const int K = plan.K;
    const uint64_t full = K >= 64 ? ~0ull : ((1ull << K) - 1ull);
    plan.mixed_mask = full;
    plan.mixed_rel_err = 0.0;
    if (plan.auto_prec != TFX_PREC_F64 || K < 2 || K > 16 || !std::isfinite(plan.probe_rel_err)) return;
    std::vector<std::pair<double, int>> gain;
    for (int k = 0; k < K; ++k) gain.push_back({probe_mixed(plan, 1ull << k, kProbeLen / 2), k});
    std::sort(gain.begin(), gain.end());  // promoting this section alone leaves the smallest error -> first
    uint64_t mask = 0;
    for (int i = 0; i < K - 1; ++i) {
        mask |= 1ull << gain[i].second;
        const double err = probe_mixed(plan, mask, kProbeLen);
        if (err <= kAutoF32Bound) {
            plan.mixed_mask = mask;
            plan.mixed_rel_err = err;
            return;
        }
    }
}

struct Cache {
    std::mutex mu;
    std::list<std::string> order;  // most recent first
    std::unordered_map<std::string, std::pair<std::shared_ptr<const SosPlan>, std::list<std::string>::iterator>> map;
    static constexpr size_t kMax = 256;
};
Cache &cache() {
    static Cache c;
    return c;
}

}  // namespace

std::shared_ptr<const SosPlan> get_sos_plan(const double *sos_host, int K) {
    if (sos_host == nullptr || K < 1 || K > TFX_SOS_MAX_K) {
        set_error("sos cascade: K must be in [1, %d] and sos non-NULL (got K=%d)", TFX_SOS_MAX_K, K);
        return nullptr;
    }
    std::string key(reinterpret_cast<const char *>(sos_host), sizeof(double) * 6 * static_cast<size_t>(K));
    Cache &c = cache();
    {
        std::lock_guard<std::mutex> lk(c.mu);
        auto it = c.map.find(key);
        if (it != c.map.end()) {
            c.order.splice(c.order.begin(), c.order, it->second.second);
            return it->second.first;
        }
    }
    auto plan = std::make_shared<SosPlan>();
    plan->K = K;
    plan->sec.resize(K);
    for (int k = 0; k < K; ++k) {
        const double *r = sos_host + 6 * k;
        for (int i = 0; i < 6; ++i)
            if (!std::isfinite(r[i])) {
                set_error("sos cascade: coefficient [%d,%d] is not finite", k, i);
                return nullptr;
            }
        plan->sec[k] = {r[0], r[1], r[2], r[4], r[5]};  // a0 (r[3]) == 1 is ignored, like the reference
    }
    for (int k0 = 0; k0 < K; k0 += TFX_SOS_MAX_FUSED) {
        SosPass p;
        p.k0 = k0;
        p.k = std::min(TFX_SOS_MAX_FUSED, K - k0);
        analyse_pass(plan->sec.data() + k0, p);
        plan->passes.push_back(p);
    }
    probe_precision(*plan);
    choose_mixed_mask(*plan);
    {
        std::lock_guard<std::mutex> lk(c.mu);
        auto it = c.map.find(key);
        if (it != c.map.end()) return it->second.first;
        c.order.push_front(key);
        c.map.emplace(key, std::make_pair(std::shared_ptr<const SosPlan>(plan), c.order.begin()));
        if (c.map.size() > Cache::kMax) {
            c.map.erase(c.order.back());
            c.order.pop_back();
        }
    }
    return plan;
}

Segmentation choose_segmentation(int64_t C, int64_t T, int64_t warm_needed, int64_t capacity, bool no_split, int oversub) {
    Segmentation g;
    g.S = 1;
    g.Lseg = T;
    g.warm = 0;
    if (no_split || warm_needed < 0 || C <= 0 || T <= 0) return g;
    // Segment starts (and warm-up starts) are kept 64-element aligned so that every 256-byte
    // chunk a stream moves is exactly two full 128-byte lines in both directions.
    constexpr int64_t kAlign = 64;
    const int64_t warm = (warm_needed + kAlign - 1) / kAlign * kAlign;
    // One full wave of streams is the minimum worth having: fewer leaves SMs idle.  A
    // segment must be at least as long as its warm-up (<= 2x work) and long enough to
    // amortise the pipeline prologue.
    int64_t S = capacity / C;
    if (S < 2) return g;
    S = std::min<int64_t>(S, T / std::max<int64_t>(512, warm));
    if (oversub > 1) {
        // Dynamic scheduling wants several items per resident warp; take them only while the
        // warm-up stays <= 1/kWarmDiv of a segment: the warm-up launch filters S * warm extra samples per channel.
        // (1/16 in round 1: the 128-channel shard each of 8 GPUs gets from config 2 then ran 7696 segments of 3742
        // samples with a 256-sample warm-up; 1/32 measured best on that shard: 5.12 -> 5.05 ms, tools/shard_time.py.)
        static const int64_t kWarmDiv = [] {
            const char *e = std::getenv("TFX_WARM_DIV");
            const int v = e ? std::atoi(e) : 0;
            return static_cast<int64_t>(v >= 4 && v <= 1024 ? v : 32);
        }();
        // ... but never settle for a whole number of items per resident warp when the warm-up is long.  With config 4's
        // mixed-precision chain (warm-up 1728 samples) the 1/kWarmDiv rule alone leaves ONE item per warp, and one or
        // exactly two items per warp -- every warp starts and ends its items together with all the others -- measured
        // 12-25 % slower than 1.7 items per warp on the same data (tools/cfg4_shard.py, 2048 / 1024 / 512 / 256 ch x 60 s:
        // 8.3-9.0 / 5.0 / 2.56 / 1.36 ms at 1.0 or 2.0 items per warp, 8.3 / 4.4 / 2.26 / 1.22 ms at 1.7-2.3).  So: at least
        // 1.73 items per warp as long as the warm-up stays <= 1/4 of a segment.
        const int64_t by_warm = T / std::max<int64_t>(4096, kWarmDiv * warm);
        // (all or nothing: 1.3 items per warp bought with a 25 % warm-up lost 5 % on the 8-branch SUM bank's 512-channel shard)
        int64_t staggered = (173 * capacity / C + 99) / 100;
        {
            const int64_t limit = T / std::max<int64_t>(4096, 4 * warm);
            staggered = staggered * 20 <= limit * 21 ? std::min(staggered, limit) : 0;  // (within 5 % of the limit still counts)
        }
        const int64_t fine = std::min<int64_t>(capacity / C * oversub, std::max(by_warm, staggered));
        S = std::max(S, fine);
    }
    while (S >= 2) {
        int64_t L = (T + S - 1) / S;
        L = (L + kAlign - 1) / kAlign * kAlign;
        const int64_t S2 = (T + L - 1) / L;
        const int64_t last = T - (S2 - 1) * L;
        if (S2 >= 2 && last >= 2) {
            g.S = S2;
            g.Lseg = L;
            g.warm = warm;
            if (std::getenv("TFX_PLAN_DEBUG")) std::fprintf(stderr, "[tfx plan] C=%lld T=%lld warm=%lld capacity=%lld -> S=%lld Lseg=%lld\n", (long long)C, (long long)T, (long long)warm, (long long)capacity, (long long)g.S, (long long)g.Lseg);
            return g;
        }
        if (S2 < 2) break;
        --S;
    }
    return g;
}

}  // namespace tfx
