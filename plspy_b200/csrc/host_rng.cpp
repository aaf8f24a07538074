// Host-side resampling index generator, bit-identical to the reference's use of numpy's legacy global RNG.
//
// The reference draws every resample with np.random.permutation / np.random.choice on the global MT19937
// stream (plspy/core/resample.py:63-79, 132-160; call order SURVEY.md App. B) -- ~80 numpy calls per permutation,
// 1.5 s for 5000 permutations of the bench design, i.e. many times the GPU time of the whole test.  These
// functions continue the SAME stream (state = the 624 key words + position of np.random.get_state()) with the
// same algorithms, so `np.random.seed(k); PLS(...)` draws the same resamples as plspy:
//   * next_uint32      : MT19937 with the standard tempering (numpy/random/src/mt19937);
//   * interval(max)    : numpy's legacy random_interval -- smallest bit mask >= max, 32-bit draws, rejection;
//   * shuffle(x, n)    : for i = n-1 .. 1: j = interval(i); swap(x[i], x[j])        (RandomState.shuffle)
//   * choice(n, n)     : n masked-rejection draws with rng = n - 1, none when n == 1  (RandomState.randint path)
// No CUDA in this file; it is part of libplsb200.so because it feeds the kernels' index matrices.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../include/plsb200.h"

namespace {

struct MT {
    uint32_t* key;
    int pos;
    inline void regen() {
        const int N = 624, M = 397;
        const uint32_t A = 0x9908b0dfu, UP = 0x80000000u, LO = 0x7fffffffu;
        int i;
        uint32_t y;
        for (i = 0; i < N - M; ++i) {
            y = (key[i] & UP) | (key[i + 1] & LO);
            key[i] = key[i + M] ^ (y >> 1) ^ ((uint32_t)(-(int32_t)(y & 1)) & A);
        }
        for (; i < N - 1; ++i) {
            y = (key[i] & UP) | (key[i + 1] & LO);
            key[i] = key[i + (M - N)] ^ (y >> 1) ^ ((uint32_t)(-(int32_t)(y & 1)) & A);
        }
        y = (key[N - 1] & UP) | (key[0] & LO);
        key[N - 1] = key[M - 1] ^ (y >> 1) ^ ((uint32_t)(-(int32_t)(y & 1)) & A);
        pos = 0;
    }
    inline uint32_t next() {
        if (pos >= 624) regen();
        uint32_t y = key[pos++];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        return y;
    }
    inline uint32_t interval(uint32_t max) {       // uniform on [0, max], numpy legacy random_interval
        if (max == 0) return 0;
        uint32_t mask = max;
        mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
        uint32_t v;
        while ((v = next() & mask) > max) {
        }
        return v;
    }
    template <class T>
    inline void shuffle(T* x, int n, int stride = 1) {
        for (int i = n - 1; i >= 1; --i) {
            const int j = (int)interval((uint32_t)i);
            const T t = x[(size_t)i * stride];
            x[(size_t)i * stride] = x[(size_t)j * stride];
            x[(size_t)j * stride] = t;
        }
    }
};

bool check_state(const uint32_t* key, const int32_t* pos) { return key != nullptr && pos != nullptr && *pos >= 0 && *pos <= 624; }

}  // namespace

// builds the base subject x condition grid (all groups concatenated); returns S (subjects), -1 for ragged designs
static int base_grid(const int32_t* cond_order, int G, int C, std::vector<int32_t>& base) {
    int S = 0;
    for (int g = 0; g < G; ++g) {
        for (int c = 1; c < C; ++c)
            if (cond_order[g * C + c] != cond_order[g * C]) return -1;
        S += cond_order[g * C];
    }
    base.resize((size_t)S * C);
    int start = 0, s0 = 0;
    for (int g = 0; g < G; ++g) {
        const int n = cond_order[g * C];
        for (int c = 0; c < C; ++c)
            for (int s = 0; s < n; ++s) base[(size_t)(s0 + s) * C + c] = start + c * n + s;
        start += n * C;
        s0 += n;
    }
    return S;
}

static void boot_draw(MT& mt, const int32_t* cond_order, int G, int C, std::vector<int32_t>& pick, int32_t* o) {
    int start = 0;
    for (int g = 0; g < G; ++g) {
        const int n = cond_order[g * C];
        pick.resize(n);
        for (int s = 0; s < n; ++s) pick[s] = (int32_t)mt.interval((uint32_t)(n - 1));
        for (int c = 0; c < C; ++c)
            for (int s = 0; s < n; ++s) o[start + c * n + s] = start + c * n + pick[s];
        start += n * C;
    }
}

// Task-method permutations (resample.py:44-79): rows are ordered group -> condition -> subject; the subject x
// condition grid of ALL groups concatenated has every subject row shuffled, then every condition column shuffled
// across all subjects; the index vector is the grid flattened condition-major.  out_task: count x N int32.
// beh_rows > 0 (multiblock, bootstrap_permutation.py:342-347): each task draw is followed by one
// np.random.permutation(beh_rows) for the behaviour block -> out_beh: count x beh_rows.
extern "C" int plsb200_host_task_permutations(uint32_t* key, int32_t* pos, const int32_t* cond_order, int G, int C,
                                              int beh_rows, int count, int32_t* out_task, int32_t* out_beh) {
    if (!check_state(key, pos) || !cond_order || !out_task || G < 1 || C < 1 || count < 0 || beh_rows < 0 ||
        (beh_rows > 0 && !out_beh))
        return PLSB200_EINVAL;
    MT mt{key, *pos};
    std::vector<int32_t> grid, base;
    const int S = base_grid(cond_order, G, C, base);
    if (S < 0) return PLSB200_EUNSUPPORTED;          // ragged designs: the caller uses the numpy path
    for (int r = 0; r < count; ++r) {
        grid = base;
        for (int s = 0; s < S; ++s) mt.shuffle(grid.data() + (size_t)s * C, C);
        int32_t* o = out_task + (size_t)r * S * C;
        for (int c = 0; c < C; ++c) {
            for (int s = 0; s < S; ++s) o[(size_t)c * S + s] = grid[(size_t)s * C + c];
            mt.shuffle(o + (size_t)c * S, S);
        }
        if (beh_rows > 0) {
            int32_t* b = out_beh + (size_t)r * beh_rows;
            for (int i = 0; i < beh_rows; ++i) b[i] = i;
            mt.shuffle(b, beh_rows);
        }
    }
    *pos = mt.pos;
    return PLSB200_OK;
}

// Bootstrap draws (resample.py:125-160): per group np.random.choice(n_g, n_g), the same subjects for every
// condition, flattened condition-major within the group.  out: count x N int32.
// cond_order2 != NULL (multiblock, bootstrap_permutation.py:545-554): each draw is followed by an independent draw
// for the behaviour block with design cond_order2 (G x C2) -> out2: count x N2.
extern "C" int plsb200_host_bootstrap_draws(uint32_t* key, int32_t* pos, const int32_t* cond_order, int G, int C,
                                            const int32_t* cond_order2, int C2, int count, int32_t* out,
                                            int32_t* out2) {
    if (!check_state(key, pos) || !cond_order || !out || G < 1 || C < 1 || count < 0 ||
        (cond_order2 && (!out2 || C2 < 1)))
        return PLSB200_EINVAL;
    MT mt{key, *pos};
    std::vector<int32_t> tmp;
    if (base_grid(cond_order, G, C, tmp) < 0) return PLSB200_EUNSUPPORTED;
    if (cond_order2 && base_grid(cond_order2, G, C2, tmp) < 0) return PLSB200_EUNSUPPORTED;
    int N = 0, N2 = 0;
    for (int g = 0; g < G; ++g) {
        N += cond_order[g * C] * C;
        if (cond_order2) N2 += cond_order2[g * C2] * C2;
    }
    std::vector<int32_t> pick;
    for (int r = 0; r < count; ++r) {
        boot_draw(mt, cond_order, G, C, pick, out + (size_t)r * N);
        if (cond_order2) boot_draw(mt, cond_order2, G, C2, pick, out2 + (size_t)r * N2);
    }
    *pos = mt.pos;
    return PLSB200_OK;
}

// np.random.permutation(n), `count` times (behaviour PLS permutations).  out: count x n int32.
extern "C" int plsb200_host_row_permutations(uint32_t* key, int32_t* pos, int n, int count, int32_t* out) {
    if (!check_state(key, pos) || !out || n < 1 || count < 0) return PLSB200_EINVAL;
    MT mt{key, *pos};
    for (int r = 0; r < count; ++r) {
        int32_t* o = out + (size_t)r * n;
        for (int i = 0; i < n; ++i) o[i] = i;
        mt.shuffle(o, n);
    }
    *pos = mt.pos;
    return PLSB200_OK;
}

// Split-half draws (split_half_resampling.py:136 / :555 and :271, :282 / :692, :703; the two routines draw the same
// sequence): `count` real splits of one np.random.permutation(n_g) per group, concatenated per split into
// out_real (count x sum n_g), THEN `count` null splits of np.random.permutation(nsub) -> out_null_subj (count x nsub)
// followed by np.random.permutation(n_rows) -> out_null_rows (count x n_rows).
extern "C" int plsb200_host_split_draws(uint32_t* key, int32_t* pos, const int32_t* group_sizes, int G, int nsub,
                                        int n_rows, int count, int32_t* out_real, int32_t* out_null_subj,
                                        int32_t* out_null_rows) {
    if (!check_state(key, pos) || !group_sizes || !out_real || !out_null_subj || !out_null_rows || G < 1 || nsub < 1 ||
        n_rows < 1 || count < 0)
        return PLSB200_EINVAL;
    int S = 0;
    for (int g = 0; g < G; ++g) {
        if (group_sizes[g] < 1) return PLSB200_EINVAL;
        S += group_sizes[g];
    }
    MT mt{key, *pos};
    for (int r = 0; r < count; ++r) {
        int32_t* o = out_real + (size_t)r * S;
        for (int g = 0; g < G; ++g) {
            const int n = group_sizes[g];
            for (int i = 0; i < n; ++i) o[i] = i;
            mt.shuffle(o, n);
            o += n;
        }
    }
    for (int r = 0; r < count; ++r) {
        int32_t* a = out_null_subj + (size_t)r * nsub;
        for (int i = 0; i < nsub; ++i) a[i] = i;
        mt.shuffle(a, nsub);
        int32_t* b = out_null_rows + (size_t)r * n_rows;
        for (int i = 0; i < n_rows; ++i) b[i] = i;
        mt.shuffle(b, n_rows);
    }
    *pos = mt.pos;
    return PLSB200_OK;
}
