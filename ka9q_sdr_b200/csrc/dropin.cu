// Drop-in layer (include/ka9q_b200.h, section A): the reference's filter.h API, same names, same struct layouts,
// same error behaviour, with every transform / multiply / design step on the GPU and host mirrors kept coherent.
//   create_filter_input  filter.c:54     execute_filter_input  filter.c:146    delete_filter_input  filter.c:254
//   create_filter_output filter.c:97     execute_filter_output filter.c:175    delete_filter_output filter.c:264
//   set_filter filter.c:500   window_filter :365   window_rfilter :420   make_kaiser :337   noise_gain :472
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../include/ka9q_b200.h"
#include "bigfft.cuh"
#include "design.cuh"
#include "fft_regs.cuh"
#include "util.cuh"

using namespace k9;

extern "C" {
float Kaiser_beta = 3.0;  // filter.c:279
}

namespace {

struct InPriv {
  struct filter_in pub;  // MUST be first: callers hold &pub
  unsigned magic;
  int device;
  int N;
  int nbins;             // N (COMPLEX) or N/2+1 (REAL)
  BigFftPlan plan;
  void* d_ring;          // N samples (float2 or float), logical index i at ring[(off+i) % N]
  long long off;
  float2 *d_fdomain, *d_tmp0, *d_tmp1;
  cudaStream_t st;
};
struct OutPriv {
  struct filter_out pub;  // MUST be first
  unsigned magic;
  int N_dec;
  int rbins;              // response bins: N_dec, or N_dec/2+1 for REAL output
  BigFftPlan plan;
  bool have_plan;
  float2 *d_resp, *d_ff, *d_out, *d_tmp0, *d_tmp1;
};
constexpr unsigned IN_MAGIC = 0x6b39494eu, OUT_MAGIC = 0x6b394f55u;

int pick_device() {
  const char* e = getenv("KA9Q_B200_DEVICE");
  return e ? atoi(e) : 0;
}

// Select + multiply for every in/out type combination of filter.c:206-249, producing a FULL N_dec-point spectrum
// ready for a complex backward transform (REAL output = Hermitian-symmetric spectrum, real part taken afterwards).
__global__ void select_multiply_kernel(const float2* __restrict__ X, int N, int in_real, const float2* __restrict__ R,
                                       int Nd, int out_type, float2* __restrict__ ff) {
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < Nd; p += gridDim.x * blockDim.x) {
    const int half = Nd / 2;
    float2 y;
    if (out_type == REAL) {
      // positive half, folded; negative half = conjugate mirror (what c2r implies)
      const int q = p <= half ? p : Nd - p;
      y = cmul(R[q], X[q]);  // filter.c:206-208
      if (!in_real && q >= 1 && q < half) {
        // y += conj(R[Nd-q] * X[N-q])   (filter.c:232-234)
        const float2 m = cmul(R[Nd - q], X[N - q]);
        y.x += m.x;
        y.y -= m.y;
      }
      if (q == 0 || 2 * q == Nd) y.y = 0.f;
      if (p > half) y.y = -y.y;
    } else {
      if (p <= half) {
        y = cmul(R[p], X[p]);
      } else if (in_real) {
        const float2 x = X[Nd - p];  // F[-f] = conj(F[+f])  (filter.c:214-216)
        y = cmul(R[p], make_float2(x.x, -x.y));
      } else {
        y = cmul(R[p], X[N - (Nd - p)]);  // filter.c:225-227
      }
    }
    ff[p] = y;
  }
}
// CROSS_CONJ butterfly (filter.c:239-249), one thread per (p, Nd-p) pair
__global__ void cross_conj_kernel(float2* __restrict__ ff, int Nd) {
  for (int p = 1 + blockIdx.x * blockDim.x + threadIdx.x; p < Nd / 2; p += gridDim.x * blockDim.x) {
    const float2 pos = ff[p], neg = ff[Nd - p];
    ff[p] = make_float2(pos.x + neg.x, pos.y - neg.y);
    ff[Nd - p] = make_float2(neg.x - pos.x, neg.y + pos.y);
  }
}

int blocks_for(int n) { return (n + 255) / 256 > 1024 ? 1024 : (n + 255) / 256; }

}  // namespace

extern "C" {

void* ka9q_alloc(size_t bytes) {
  void* p = nullptr;
  if (posix_memalign(&p, 64, bytes ? bytes : 64) != 0) return nullptr;
  return p;
}
void ka9q_free(void* p) { free(p); }

int make_kaiser(float* const window, unsigned int const M, float const beta) {
  if (window == NULL) return -1;
  kaiser_window_host(window, M, beta);
  return 0;
}

struct filter_in* create_filter_input(unsigned int const L, unsigned int const M, enum filtertype const in_type) {
  int const N = L + M - 1;
  if (ka9q_device_count() <= pick_device()) {
    set_error("create_filter_input: no CUDA device (this library has no CPU fallback)");
    fprintf(stderr, "ka9q_b200: %s\n", get_error());
    return NULL;
  }
  InPriv* m = (InPriv*)calloc(1, sizeof(InPriv));
  if (!m) return NULL;
  m->magic = IN_MAGIC;
  m->device = pick_device();
  m->N = N;
  cudaSetDevice(m->device);
  pthread_mutex_init(&m->pub.filter_mutex, NULL);
  m->pub.blocknum = 0;
  pthread_cond_init(&m->pub.filter_cond, NULL);
  enum filtertype t = in_type;
  if (t != COMPLEX && t != REAL) {
    fprintf(stderr, "Filter input type %d, assuming complex\n", in_type);  // filter.c:68
    t = COMPLEX;
  }
  m->pub.in_type = in_type;
  m->pub.ilen = L;
  m->pub.impulse_length = M;
  if (bigfft_plan_create(&m->plan, N) != 0) {
    set_error("create_filter_input: FFT size %d has no supported factorisation", N);
    fprintf(stderr, "ka9q_b200: %s\n", get_error());
    pthread_mutex_destroy(&m->pub.filter_mutex);
    pthread_cond_destroy(&m->pub.filter_cond);
    free(m);
    return NULL;
  }
  const size_t esz = (t == REAL) ? sizeof(float) : sizeof(float2);
  m->nbins = (t == REAL) ? N / 2 + 1 : N;
  m->pub.fdomain = ka9q_alloc(sizeof(float2) * m->nbins);
  m->pub.input_buffer.r = (float*)ka9q_alloc(esz * N);
  memset(m->pub.input_buffer.r, 0, esz * (M - 1));  // clear earlier state (filter.c:77,87)
  m->pub.input.r = (float*)((char*)m->pub.input_buffer.r + esz * (M - 1));
  m->pub.fwd_plan = m;
  bool ok = cudaStreamCreateWithFlags(&m->st, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaMalloc(&m->d_ring, esz * N) == cudaSuccess;
  ok = ok && cudaMemset(m->d_ring, 0, esz * N) == cudaSuccess;
  ok = ok && cudaMalloc(&m->d_fdomain, sizeof(float2) * N) == cudaSuccess;
  ok = ok && cudaMalloc(&m->d_tmp0, sizeof(float2) * N) == cudaSuccess;
  ok = ok && cudaMalloc(&m->d_tmp1, sizeof(float2) * N) == cudaSuccess;
  if (!ok) {
    set_error("create_filter_input: device allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
    fprintf(stderr, "ka9q_b200: %s\n", get_error());
    delete_filter_input(&m->pub);  // frees the plan, the host mirrors and whatever device buffers were obtained
    return NULL;
  }
  return &m->pub;
}

struct filter_out* create_filter_output(struct filter_in* master, void* response, unsigned int decimate,
                                        enum filtertype out_type) {
  if (master == NULL) return NULL;
  InPriv* mp = (InPriv*)master;
  int const N = master->ilen + master->impulse_length - 1;
  int const N_dec = N / decimate;
  if ((N % decimate) != 0)
    fprintf(stderr, "Warning: FFT size %'u is not divisible by decimation ratio %d\n", N, decimate);  // filter.c:106-107
  OutPriv* s = (OutPriv*)calloc(1, sizeof(OutPriv));
  if (s == NULL) return NULL;
  s->magic = OUT_MAGIC;
  s->N_dec = N_dec;
  cudaSetDevice(mp->device);
  s->pub.master = master;
  s->pub.out_type = out_type;
  s->pub.decimate = decimate;
  s->pub.olen = master->ilen / decimate;
  s->pub.response = response;
  pthread_mutex_init(&s->pub.response_mutex, NULL);
  if (response != NULL)
    s->pub.noise_gain = noise_gain(&s->pub);
  else
    s->pub.noise_gain = NAN;
  s->rbins = (out_type == REAL) ? N_dec / 2 + 1 : N_dec;
  if (bigfft_plan_create(&s->plan, N_dec) != 0) {
    set_error("create_filter_output: inverse FFT size %d has no supported factorisation", N_dec);
    fprintf(stderr, "ka9q_b200: %s\n", get_error());
    free(s);
    return NULL;
  }
  s->have_plan = true;
  if (out_type == REAL) {
    s->pub.f_fdomain = ka9q_alloc(sizeof(float2) * (N_dec / 2 + 1));
    s->pub.output_buffer.r = (float*)ka9q_alloc(sizeof(float) * N_dec);
    s->pub.output.r = s->pub.output_buffer.r + N_dec - s->pub.olen;  // filter.c:140
  } else {
    s->pub.f_fdomain = ka9q_alloc(sizeof(float2) * N_dec);
    s->pub.output_buffer.r = (float*)ka9q_alloc(sizeof(float2) * N_dec);
    s->pub.output.r = (float*)((float2*)s->pub.output_buffer.r + N_dec - s->pub.olen);  // filter.c:131
  }
  s->pub.rev_plan = s;
  bool ok = cudaMalloc(&s->d_resp, sizeof(float2) * N_dec) == cudaSuccess;
  ok = ok && cudaMalloc(&s->d_ff, sizeof(float2) * N_dec) == cudaSuccess;
  ok = ok && cudaMalloc(&s->d_out, sizeof(float2) * N_dec) == cudaSuccess;
  ok = ok && cudaMalloc(&s->d_tmp0, sizeof(float2) * N_dec) == cudaSuccess;
  ok = ok && cudaMalloc(&s->d_tmp1, sizeof(float2) * N_dec) == cudaSuccess;
  if (!ok) {
    set_error("create_filter_output: device allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
    s->pub.response = NULL;  // the caller keeps its response on failure
    delete_filter_output(&s->pub);
    return NULL;
  }
  return &s->pub;
}

int execute_filter_input(struct filter_in* const master) {
  if (master == NULL) return -1;
  InPriv* m = (InPriv*)master;
  cudaSetDevice(m->device);
  const bool real = master->in_type == REAL;
  const size_t esz = real ? sizeof(float) : sizeof(float2);
  const int N = m->N, L = master->ilen, M = master->impulse_length;
  // upload the L new samples behind the M-1 samples of history already on the device
  long long pos = (m->off + (M - 1)) % N;
  long long done = 0;
  while (done < L) {
    long long chunk = (L - done) < (N - pos) ? (L - done) : (N - pos);
    if (cudaMemcpyAsync((char*)m->d_ring + pos * esz, (const char*)master->input.r + done * esz, chunk * esz,
                        cudaMemcpyHostToDevice, m->st) != cudaSuccess)
      return -1;
    done += chunk;
    pos = (pos + chunk) % N;
  }
  BigFftIn in;
  in.in_mode = real ? IN_RING_R32 : IN_RING_C32;
  in.in = m->d_ring;
  in.ring_cap = N;
  in.ring_off = m->off;
  in.ring_step = 0;
  if (bigfft_exec(&m->plan, in, m->d_fdomain, N, m->d_tmp0, m->d_tmp1, 1, -1, m->st)) return -1;  // filter.c:151
  if (cudaMemcpyAsync(master->fdomain, m->d_fdomain, sizeof(float2) * m->nbins, cudaMemcpyDeviceToHost, m->st) != cudaSuccess)
    return -1;
  if (cudaStreamSynchronize(m->st) != cudaSuccess) {
    set_error("execute_filter_input: %s", cudaGetErrorString(cudaGetLastError()));
    return -1;
  }
  m->off = (m->off + L) % N;  // device-side overlap-save: the ring just advances
  // notify slaves (filter.c:154-157)
  pthread_mutex_lock(&master->filter_mutex);
  master->blocknum++;
  pthread_cond_broadcast(&master->filter_cond);
  pthread_mutex_unlock(&master->filter_mutex);
  // host-side overlap-save so the caller-visible buffer behaves as in the reference (filter.c:159-170)
  memmove(master->input_buffer.r, (char*)master->input_buffer.r + (size_t)L * esz, (size_t)(M - 1) * esz);
  return 0;
}

int execute_filter_output(struct filter_out* const slave) {
  if (slave == NULL) return -1;
  OutPriv* s = (OutPriv*)slave;
  struct filter_in* master = slave->master;
  InPriv* m = (InPriv*)master;
  int const N = master->ilen + master->impulse_length - 1;
  int const N_dec = N / slave->decimate;
  // wait for a new block (filter.c:195-199)
  pthread_mutex_lock(&master->filter_mutex);
  while (slave->blocknum == master->blocknum) pthread_cond_wait(&master->filter_cond, &master->filter_mutex);
  slave->blocknum = master->blocknum;
  pthread_mutex_unlock(&master->filter_mutex);
  cudaSetDevice(m->device);
  cudaStream_t st = m->st;
  const int out_type = slave->out_type;  // may be rewritten by the caller between blocks (linear.c:117-120)
  // bins the multiply reads: REAL->REAL only touches the positive half (filter.c:206-208); every other combination
  // reads response[N_dec/2+1 .. N_dec) too (filter.c:214-216,225-227,232-234)
  const int rbins = (out_type == REAL && master->in_type == REAL) ? N_dec / 2 + 1 : N_dec;
  pthread_mutex_lock(&slave->response_mutex);
  if (slave->response == NULL) {
    pthread_mutex_unlock(&slave->response_mutex);
    set_error("execute_filter_output: response not set");
    return -1;
  }
  cudaError_t e = cudaMemcpyAsync(s->d_resp, slave->response, sizeof(float2) * rbins, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  pthread_mutex_unlock(&slave->response_mutex);
  if (e != cudaSuccess) return -1;
  select_multiply_kernel<<<blocks_for(N_dec), 256, 0, st>>>(m->d_fdomain, N, master->in_type == REAL ? 1 : 0, s->d_resp,
                                                            N_dec, out_type, s->d_ff);
  if (out_type == CROSS_CONJ) cross_conj_kernel<<<blocks_for(N_dec / 2), 256, 0, st>>>(s->d_ff, N_dec);
  BigFftIn in;
  in.in = s->d_ff;
  in.in_batch_stride = N_dec;
  if (bigfft_exec(&s->plan, in, s->d_out, N_dec, s->d_tmp0, s->d_tmp1, 1, +1, st)) return -1;  // filter.c:250
  if (out_type == REAL)
    e = cudaMemcpy2DAsync(slave->output_buffer.r, sizeof(float), s->d_out, sizeof(float2), sizeof(float), N_dec,
                          cudaMemcpyDeviceToHost, st);
  else
    e = cudaMemcpyAsync(slave->output_buffer.r, s->d_out, sizeof(float2) * N_dec, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) {
    set_error("execute_filter_output: %s", cudaGetErrorString(e));
    return -1;
  }
  return 0;
}

int delete_filter_input(struct filter_in* const master) {
  if (master == NULL) return 0;
  InPriv* m = (InPriv*)master;
  cudaSetDevice(m->device);
  if (m->st) cudaStreamSynchronize(m->st);
  bigfft_plan_destroy(&m->plan);
  cudaFree(m->d_ring);
  cudaFree(m->d_fdomain);
  cudaFree(m->d_tmp0);
  cudaFree(m->d_tmp1);
  if (m->st) cudaStreamDestroy(m->st);
  pthread_mutex_destroy(&master->filter_mutex);
  pthread_cond_destroy(&master->filter_cond);
  ka9q_free(master->input_buffer.r);
  ka9q_free(master->fdomain);
  free(m);
  return 0;
}

int delete_filter_output(struct filter_out* const slave) {
  if (slave == NULL) return 0;
  OutPriv* s = (OutPriv*)slave;
  pthread_mutex_destroy(&slave->response_mutex);
  if (s->have_plan) bigfft_plan_destroy(&s->plan);
  cudaFree(s->d_resp);
  cudaFree(s->d_ff);
  cudaFree(s->d_out);
  cudaFree(s->d_tmp0);
  cudaFree(s->d_tmp1);
  ka9q_free(slave->output_buffer.r);
  ka9q_free(slave->response);  // owned by the slave (filter.c:271)
  ka9q_free(slave->f_fdomain);
  free(s);
  return 0;
}

float noise_gain(struct filter_out const* const filter) {
  if (filter == NULL) return NAN;
  struct filter_in* master = filter->master;
  int const N = master->ilen + master->impulse_length - 1;
  int const N_dec = N / filter->decimate;
  const float2* r = (const float2*)filter->response;
  float sum = 0;
  int const bins = (master->in_type == REAL && filter->out_type == REAL) ? N_dec / 2 + 1 : N_dec;
  for (int i = 0; i < bins; i++) sum += r[i].x * r[i].x + r[i].y * r[i].y;
  if (filter->out_type == REAL || filter->out_type == CROSS_CONJ)
    return 2 * N * sum;
  else
    return N * sum;
}

static int window_common(int const L, int const M, void* const response, float const beta, bool real) {
  if (response == NULL) return -1;
  int const N = L + M - 1;
  if (ka9q_device_count() <= pick_device()) {
    set_error("window_filter: no CUDA device (no CPU fallback)");
    return -1;
  }
  cudaSetDevice(pick_device());
  BigFftPlan plan;
  if (bigfft_plan_create(&plan, N) != 0 || plan.npass >= 3) {
    set_error("window_filter: size %d unsupported", N);
    return -1;
  }
  std::vector<float> w(M);
  kaiser_window_host(w.data(), M, beta);
  float2 *d_resp = nullptr, *d_work = nullptr, *d_half = nullptr;
  float* d_w = nullptr;
  const int nh = N / 2 + 1;
  bool ok = cudaMalloc(&d_resp, sizeof(float2) * N) == cudaSuccess && cudaMalloc(&d_work, sizeof(float2) * 2 * N) == cudaSuccess &&
            cudaMalloc(&d_w, sizeof(float) * M) == cudaSuccess && cudaMalloc(&d_half, sizeof(float2) * nh) == cudaSuccess;
  int r = ok ? 0 : -1;
  if (ok) {
    cudaMemcpy(d_w, w.data(), sizeof(float) * M, cudaMemcpyHostToDevice);
    if (real) {
      cudaMemcpy(d_half, response, sizeof(float2) * nh, cudaMemcpyHostToDevice);
      r = window_rfilter_device(&plan, M, d_half, d_resp, 1, d_w, d_work, 0);
      if (r == 0 && cudaMemcpy(response, d_resp, sizeof(float2) * nh, cudaMemcpyDeviceToHost) != cudaSuccess) r = -1;
    } else {
      cudaMemcpy(d_resp, response, sizeof(float2) * N, cudaMemcpyHostToDevice);
      r = window_filter_device(&plan, M, d_resp, 1, d_w, d_work, 0);
      if (r == 0 && cudaMemcpy(response, d_resp, sizeof(float2) * N, cudaMemcpyDeviceToHost) != cudaSuccess) r = -1;
    }
  }
  cudaFree(d_resp);
  cudaFree(d_work);
  cudaFree(d_w);
  cudaFree(d_half);
  bigfft_plan_destroy(&plan);
  return r;
}

int window_filter(int const L, int const M, void* const response, float const beta) {
  return window_common(L, M, response, beta, false);
}
int window_rfilter(int const L, int const M, void* const response, float const beta) {
  return window_common(L, M, response, beta, true);
}

int set_filter(struct filter_out* const slave, float const low, float const high, float const kaiser_beta) {
  if (slave == NULL) return -1;
  if (isnan(low) || isnan(high)) return -1;  // filter.c:504-505
  struct filter_in* master = slave->master;
  int const L_dec = slave->olen;
  int const M_dec = (master->impulse_length - 1) / slave->decimate + 1;
  int const N_dec = L_dec + M_dec - 1;
  int const N = master->ilen + master->impulse_length - 1;
  float gain = 1. / ((float)N);
  if (slave->out_type == REAL || slave->out_type == CROSS_CONJ) gain *= M_SQRT1_2;
  float2* response = (float2*)ka9q_alloc(sizeof(float2) * N_dec);
  if (!response) return -1;
  for (int n = 0; n < N_dec; n++) {
    float f;
    if (n <= N_dec / 2)
      f = (float)n / N_dec;
    else
      f = (float)(n - N_dec) / N_dec;
    response[n] = (f >= low && f <= high) ? make_float2(gain, 0.f) : make_float2(0.f, 0.f);
  }
  if (window_filter(L_dec, M_dec, response, kaiser_beta) != 0) {
    ka9q_free(response);
    return -1;
  }
  pthread_mutex_lock(&slave->response_mutex);
  void* tmp = slave->response;
  slave->response = response;
  slave->noise_gain = noise_gain(slave);
  pthread_mutex_unlock(&slave->response_mutex);
  ka9q_free(tmp);
  return 0;
}

}  // extern "C"
