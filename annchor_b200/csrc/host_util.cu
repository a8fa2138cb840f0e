// host_util.cu -- host-side helpers of the C ABI that are not kernels.
#include "common.cuh"

namespace annb {

// numba's np.random inside @njit as used by the sampler (annchor/utils.py:555-557,572):
// MT19937 seeded by init_genrand, np.random.shuffle = Fisher-Yates from the end with
// randint(i+1) drawn by bit-mask rejection (numba/cpython/randomimpl.py).
struct Mt {
    uint32_t mt[624];
    int idx;
    void seed(uint32_t s)
    {
        mt[0] = s;
        for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
        idx = 624;
    }
    uint32_t next()
    {
        if (idx >= 624) {
            for (int k = 0; k < 624; ++k) {
                uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
                mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            idx = 0;
        }
        uint32_t y = mt[idx++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }
    int64_t below(int64_t n)
    {
        if (n == 1) return 0;
        const int nbits = 64 - __builtin_clzll((uint64_t)(n - 1));
        for (;;) {
            int64_t r;
            if (nbits <= 32) {
                r = (int64_t)(next() & (0xffffffffu >> (32 - nbits)));
            } else {
                const uint64_t hi = next() & (0xffffffffu >> (64 - nbits));
                const uint64_t lo = next();
                r = (int64_t)((hi << 32) | lo);
            }
            if (r < n) return r;
        }
    }
};

}  // namespace annb

using namespace annb;

ANNB_API int annb_numba_rng_new(uint32_t seed, void **state)
{
    ANNB_REQUIRE(state != nullptr, ANNB_EINVAL, "state is NULL");
    Mt *m = new Mt();
    m->seed(seed);
    *state = m;
    return ANNB_OK;
}

ANNB_API int annb_numba_rng_free(void *state)
{
    delete static_cast<Mt *>(state);
    return ANNB_OK;
}

ANNB_API int annb_numba_rng_shuffle(void *state, int64_t *x, int64_t n)
{
    ANNB_REQUIRE(state && (x || n == 0), ANNB_EINVAL, "NULL argument");
    Mt *m = static_cast<Mt *>(state);
    for (int64_t i = n - 1; i > 0; --i) {
        const int64_t j = m->below(i + 1);
        const int64_t t = x[i];
        x[i] = x[j];
        x[j] = t;
    }
    return ANNB_OK;
}
