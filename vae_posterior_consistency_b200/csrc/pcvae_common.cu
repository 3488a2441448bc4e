// Error buffer, flat-parameter layout and device check shared by the C ABI entry points.
#include "pcvae_internal.cuh"

namespace pcvae {

static thread_local char g_err[512] = "";

char* err_buf() { return g_err; }

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

bool make_layout(const pcvae_model* m, Layout* L) {
    if (!m) { fail(PCVAE_EINVAL, "null model"); return false; }
    if (m->family != PCVAE_FAMILY_MLP && m->family != PCVAE_FAMILY_PNP && m->family != PCVAE_FAMILY_MLP_MASK) { fail(PCVAE_EINVAL, "unknown family %d", m->family); return false; }
    if (m->obs_dim < 1 || m->obs_dim > MAX_D) { fail(PCVAE_EINVAL, "obs_dim %d outside [1,%d]", m->obs_dim, MAX_D); return false; }
    if (m->latent_dim != LAT) { fail(PCVAE_EINVAL, "latent_dim must be %d (got %d)", LAT, m->latent_dim); return false; }
    const int D = m->obs_dim;
    int K = 0, o = 0;
    L->aug = m->family == PCVAE_FAMILY_MLP_MASK ? 1 : 0;
    L->fam = L->aug ? PCVAE_FAMILY_MLP : m->family; L->D = D;
    L->E = L->bE = L->We = L->be = 0;
    if (m->family == PCVAE_FAMILY_PNP) {
        K = m->emb_dim;
        if (K < 1 || K > MAX_K) { fail(PCVAE_EINVAL, "emb_dim %d outside [1,%d]", K, MAX_K); return false; }
        L->E = o; o += D * K;
        L->bE = o; o += D;
        L->We = o; o += K * (K + 2);
        L->be = o; o += K;
    }
    L->K = K;
    const int in1 = (m->family == PCVAE_FAMILY_PNP) ? K : (L->aug ? 2 * D : D);
    L->W1 = o; o += H1 * in1;
    L->b1 = o; o += H1;
    L->W2 = o; o += H2 * H1;
    L->b2 = o; o += H2;
    L->W3 = o; o += LAT2 * H2;
    L->b3 = o; o += LAT2;
    L->W4 = o; o += G1 * LAT;
    L->b4 = o; o += G1;
    L->W5 = o; o += G2 * G1;
    L->b5 = o; o += G2;
    L->W6 = o; o += D * G2;
    L->b6 = o; o += D;
    L->total = o;
    return true;
}

// Optional per-kernel timing hooks (pcvae_profile_events): the tensor-core launchers call prof_mark() before their
// first kernel and after every kernel; while armed, mark k records the caller's k-th CUDA event on the stream.
static thread_local cudaEvent_t* g_prof_ev = nullptr;
static thread_local int g_prof_n = 0, g_prof_i = 0;

bool pdl_enabled() {
    static const bool on = [] { const char* e = getenv("PCVAE_PDL"); return !(e && e[0] == '0'); }();
    return on;
}

void prof_mark(cudaStream_t st) {
    if (g_prof_ev && g_prof_i < g_prof_n) cudaEventRecord(g_prof_ev[g_prof_i++], st);
}

// Per-device status word of the tensor-core kernels: one int in pinned, mapped host memory.  A kernel whose bounded
// mbarrier wait runs out (tc::mbar_wait) stores a non-zero code there; the NEXT entry into the library on that device
// (device_ok, the first thing every entry point does) reports it as PCVAE_ECUDA instead of letting a corrupted
// training run continue with rc = 0.  Reading the word needs no synchronisation.
constexpr int MAX_DEVICES = 64;
static int* g_status_host[MAX_DEVICES];
static int* g_status_dev[MAX_DEVICES];

static void status_init(int dev) {
    if (dev < 0 || dev >= MAX_DEVICES || g_status_host[dev]) return;
    int* h = nullptr;
    if (cudaHostAlloc(reinterpret_cast<void**>(&h), 64, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return; }
    *h = 0;
    int* d = nullptr;
    if (cudaHostGetDevicePointer(reinterpret_cast<void**>(&d), h, 0) != cudaSuccess) { cudaGetLastError(); cudaFreeHost(h); return; }
    g_status_dev[dev] = d;
    g_status_host[dev] = h;
}

int* tc_status_ptr() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEVICES) return nullptr;
    return g_status_dev[dev];
}

int device_ok(int* n_sm) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return fail(PCVAE_EDEVICE, "no CUDA device: %s (this library has no CPU fallback)", cudaGetErrorString(e));
    static int cached_dev = -1, cached_sm = 0, cached_major = 0;
    if (dev != cached_dev) {
        int major = 0, sm = 0;
        cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
        cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
        cached_major = major; cached_sm = sm; cached_dev = dev;
    }
    if (cached_major != 10) return fail(PCVAE_EDEVICE, "device %d is compute capability %d.x; this library is sm_100a only", dev, cached_major);
    if (n_sm) *n_sm = cached_sm;
    if (dev >= 0 && dev < MAX_DEVICES) {
        if (!g_status_host[dev]) status_init(dev);
        if (g_status_host[dev]) {
            const int code = *reinterpret_cast<volatile int*>(g_status_host[dev]);
            if (code != 0) {
                *reinterpret_cast<volatile int*>(g_status_host[dev]) = 0;
                return fail(PCVAE_ECUDA, "a tensor-core kernel of an earlier call on device %d gave up waiting for an mbarrier "
                            "(kernel code %d): the results of that call are invalid", dev, code);
            }
        }
    }
    return PCVAE_OK;
}

}  // namespace pcvae

extern "C" int pcvae_profile_events(void** events, int n) {
    pcvae::g_prof_ev = reinterpret_cast<cudaEvent_t*>(events);
    pcvae::g_prof_n = events ? n : 0;
    const int used = pcvae::g_prof_i;
    pcvae::g_prof_i = 0;
    return used;
}
