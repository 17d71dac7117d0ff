// Host receive / send plumbing around the device path (BASELINE north_star: "a pinned ring buffer fed by the
// multicast receive thread feeds CUDA streams"). Plain C with pthreads; the only CUDA call is the page-locked allocation
// (through ka9q_host_alloc) so that ka9q_stream_push can DMA straight out of the ring.
//
//   receive thread   rtp_recv (main.c:288-365): recvfrom -> header checks -> packet inserted into a queue SORTED BY
//                    SEQUENCE NUMBER (main.c:347-361) -> consumer woken
//   ingest thread    the packet head of proc_samples (radio.c:51-100): pop the queue head, rtp_process, zero-fill lost
//                    samples, append to the sample ring (ka9q_ingest_datagram in rtp_glue.c does the per-packet work)
//   block ring       page-locked, a whole number of 20 ms blocks; the consumer takes whole blocks (contiguous by
//                    construction) and hands them to ka9q_stream_push
//   egress           one sendmmsg per batch of PCM packets instead of one send() per packet (audio.c:73,122): at cfg5 a
//                    GPU produces ~819 k packets/s
#define _GNU_SOURCE 1
#include <errno.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <sys/socket.h>
#include <sys/types.h>
#include <time.h>
#include <unistd.h>
#include "../../include/ka9q_b200.h"

#define RX_PKT_MAX 9000  // jumbo-frame payload; the reference's struct packet holds 8192 + headers (multicast.h)

struct rx_packet {
  struct rx_packet *next;
  uint16_t seq;
  int size;
  unsigned char content[RX_PKT_MAX];
};

struct ka9q_rx {
  ka9q_ingest ingest;
  int bytes_per_sample;
  long long block_samples, ring_blocks, ring_samples;
  unsigned char *ring;          // page-locked
  long long wr;                 // samples appended since start
  long long rd;                 // samples consumed since start (whole blocks)
  struct rx_packet *queue;      // sorted by sequence number (main.c:347-361)
  struct rx_packet *free_list;
  long long queued, reordered;  // statistics
  pthread_mutex_t qmutex;       // demod->input.qmutex
  pthread_cond_t qcond;         // demod->input.qcond
  pthread_mutex_t rmutex;       // ring indices
  pthread_cond_t rcond;         // "a block completed" / "space freed"
  pthread_t rx_thread, in_thread;
  int fd, running, threads_started;
  int pinned;                   // the ring is page-locked (cudaHostAlloc); 0 where no CUDA device exists (host-logic tests)
  unsigned char *scratch;       // one datagram's worth of samples incl. zero fill (<= 192000 + packet)
};

ka9q_rx *ka9q_rx_create(int iq_format, long long block_samples, int ring_blocks) {
  if ((iq_format != KA9Q_IQ_S16 && iq_format != KA9Q_IQ_S8) || block_samples <= 0 || ring_blocks < 2) return NULL;
  ka9q_rx *rx = calloc(1, sizeof(*rx));
  if (!rx) return NULL;
  ka9q_ingest_init(&rx->ingest, iq_format);
  rx->bytes_per_sample = iq_format == KA9Q_IQ_S16 ? 4 : 2;
  rx->block_samples = block_samples;
  rx->ring_blocks = ring_blocks;
  rx->ring_samples = block_samples * ring_blocks;
  // page-locked so that ka9q_stream_push DMAs straight out of the ring; plain aligned memory where there is no CUDA device
  // (this is host plumbing, no compute: the CPU test suite drives the queue / ring logic without a GPU)
  rx->ring = ka9q_device_count() > 0 ? ka9q_host_alloc((size_t)rx->ring_samples * rx->bytes_per_sample) : NULL;
  rx->pinned = rx->ring != NULL;
  if (!rx->ring) rx->ring = ka9q_alloc((size_t)rx->ring_samples * rx->bytes_per_sample);
  rx->scratch = malloc((size_t)(192000 + RX_PKT_MAX) * 4);
  if (!rx->ring || !rx->scratch) {
    if (rx->ring) (rx->pinned ? ka9q_host_free : ka9q_free)(rx->ring);
    free(rx->scratch);
    free(rx);
    return NULL;
  }
  pthread_mutex_init(&rx->qmutex, NULL);
  pthread_cond_init(&rx->qcond, NULL);
  pthread_mutex_init(&rx->rmutex, NULL);
  pthread_cond_init(&rx->rcond, NULL);
  rx->fd = -1;
  return rx;
}

// What rtp_recv does with one received datagram (main.c:319-361): size / payload-type checks, then the sorted insert.
// Returns 0 if queued, -1 if ignored.
int ka9q_rx_inject(ka9q_rx *rx, const void *datagram, int size) {
  if (!rx || !datagram || size > RX_PKT_MAX) return -1;
  if (size < KA9Q_RTP_MIN_SIZE) return -1;  // main.c:319-320
  const unsigned char *dp = datagram;
  const int type = dp[1] & 0x7f;
  if (type != KA9Q_IQ_PT && type != KA9Q_IQ_PT8) return -1;  // main.c:331-332
  pthread_mutex_lock(&rx->qmutex);
  struct rx_packet *pkt = rx->free_list;
  if (pkt) rx->free_list = pkt->next;
  pthread_mutex_unlock(&rx->qmutex);
  if (!pkt) pkt = malloc(sizeof(*pkt));
  if (!pkt) return -1;
  memcpy(pkt->content, datagram, (size_t)size);
  pkt->size = size;
  pkt->seq = (uint16_t)(dp[2] << 8 | dp[3]);
  // Insert onto queue sorted by sequence number, wake up the consumer (main.c:347-361; the comparison is the
  // reference's plain `>=` on the 16-bit numbers, wrap-around included)
  struct rx_packet *q_prev = NULL, *qe = NULL;
  pthread_mutex_lock(&rx->qmutex);
  for (qe = rx->queue; qe && pkt->seq >= qe->seq; q_prev = qe, qe = qe->next)
    ;
  if (qe) rx->reordered++;
  pkt->next = qe;
  if (q_prev)
    q_prev->next = pkt;
  else
    rx->queue = pkt;
  rx->queued++;
  pthread_cond_signal(&rx->qcond);
  pthread_mutex_unlock(&rx->qmutex);
  return 0;
}

// appends n samples to the ring (blocks while the ring is full and the threads are running; otherwise returns -2)
static int ring_append(ka9q_rx *rx, const unsigned char *src, long long n) {
  const int bps = rx->bytes_per_sample;
  while (n > 0) {
    pthread_mutex_lock(&rx->rmutex);
    while (rx->wr - rx->rd >= rx->ring_samples) {
      if (!rx->running) {
        pthread_mutex_unlock(&rx->rmutex);
        return -2;
      }
      pthread_cond_wait(&rx->rcond, &rx->rmutex);
    }
    long long room = rx->ring_samples - (rx->wr - rx->rd);
    const long long pos = rx->wr % rx->ring_samples;
    long long chunk = n < room ? n : room;
    if (chunk > rx->ring_samples - pos) chunk = rx->ring_samples - pos;
    pthread_mutex_unlock(&rx->rmutex);
    memcpy(rx->ring + pos * bps, src, (size_t)chunk * bps);
    pthread_mutex_lock(&rx->rmutex);
    rx->wr += chunk;
    pthread_cond_broadcast(&rx->rcond);
    pthread_mutex_unlock(&rx->rmutex);
    src += chunk * bps;
    n -= chunk;
  }
  return 0;
}

// What proc_samples does with the queue (radio.c:51-100), for every packet queued right now, in queue order.
// Returns the number of samples appended to the ring (zero fill included), or -2 if the ring is full.
long long ka9q_rx_drain(ka9q_rx *rx) {
  if (!rx) return -1;
  long long total = 0;
  for (;;) {
    pthread_mutex_lock(&rx->qmutex);
    struct rx_packet *pkt = rx->queue;
    if (pkt) rx->queue = pkt->next;
    pthread_mutex_unlock(&rx->qmutex);
    if (!pkt) break;
    const long long n = ka9q_ingest_datagram(&rx->ingest, pkt->content, pkt->size, rx->scratch, 192000 + RX_PKT_MAX);
    pthread_mutex_lock(&rx->qmutex);
    pkt->next = rx->free_list;
    rx->free_list = pkt;
    pthread_mutex_unlock(&rx->qmutex);
    if (n > 0) {
      if (ring_append(rx, rx->scratch, n)) return -2;
      total += n;
    }
  }
  return total;
}

// Pointer to `nblocks` whole, contiguous blocks at the read position, or NULL if fewer are complete (wait_ms > 0: wait
// that long for them). The memory is page-locked: pass it to ka9q_stream_push, then ka9q_rx_consume.
const void *ka9q_rx_peek_blocks(ka9q_rx *rx, int nblocks, int wait_ms) {
  if (!rx || nblocks < 1 || nblocks > rx->ring_blocks) return NULL;
  const long long need = nblocks * rx->block_samples;
  pthread_mutex_lock(&rx->rmutex);
  const long long pos = rx->rd % rx->ring_samples;
  if (pos + need > rx->ring_samples) {  // would wrap: the caller asks for fewer blocks (rd only moves by whole blocks)
    pthread_mutex_unlock(&rx->rmutex);
    return NULL;
  }
  if (wait_ms > 0 && rx->wr - rx->rd < need) {
    struct timespec ts;
    clock_gettime(CLOCK_REALTIME, &ts);
    ts.tv_sec += wait_ms / 1000;
    ts.tv_nsec += (long)(wait_ms % 1000) * 1000000L;
    if (ts.tv_nsec >= 1000000000L) {
      ts.tv_sec++;
      ts.tv_nsec -= 1000000000L;
    }
    while (rx->wr - rx->rd < need)
      if (pthread_cond_timedwait(&rx->rcond, &rx->rmutex, &ts) == ETIMEDOUT) break;
  }
  const int ok = rx->wr - rx->rd >= need;
  pthread_mutex_unlock(&rx->rmutex);
  return ok ? rx->ring + pos * rx->bytes_per_sample : NULL;
}

int ka9q_rx_consume(ka9q_rx *rx, int nblocks) {
  if (!rx || nblocks < 1) return -1;
  pthread_mutex_lock(&rx->rmutex);
  const long long need = nblocks * rx->block_samples;
  if (rx->wr - rx->rd < need) {
    pthread_mutex_unlock(&rx->rmutex);
    return -1;
  }
  rx->rd += need;
  pthread_cond_broadcast(&rx->rcond);
  pthread_mutex_unlock(&rx->rmutex);
  return 0;
}

long long ka9q_rx_blocks_ready(ka9q_rx *rx) {
  if (!rx) return -1;
  pthread_mutex_lock(&rx->rmutex);
  const long long n = (rx->wr - rx->rd) / rx->block_samples;
  pthread_mutex_unlock(&rx->rmutex);
  return n;
}

void ka9q_rx_get_stats(ka9q_rx *rx, ka9q_rx_stats *st) {
  if (!rx || !st) return;
  memset(st, 0, sizeof(*st));
  pthread_mutex_lock(&rx->qmutex);
  st->datagrams_queued = rx->queued;
  st->inserted_out_of_order = rx->reordered;
  pthread_mutex_unlock(&rx->qmutex);
  st->samples = rx->ingest.samples;
  st->zero_filled = rx->ingest.zero_filled;
  st->ignored = rx->ingest.ignored;
  st->rtp_drops = rx->ingest.rtp.drops;
  st->rtp_dupes = rx->ingest.rtp.dupes;
  st->pinned = rx->pinned;
}

static void *rx_thread_main(void *arg) {  // rtp_recv (main.c:288-365)
  ka9q_rx *rx = arg;
  unsigned char buf[RX_PKT_MAX];
  while (rx->running) {
    const ssize_t size = recv(rx->fd, buf, sizeof(buf), 0);
    if (size <= 0) {
      if (!rx->running) break;
      if (size < 0 && (errno == EAGAIN || errno == EWOULDBLOCK || errno == EINTR)) continue;
      usleep(50000);  // main.c:315-318
      continue;
    }
    ka9q_rx_inject(rx, buf, (int)size);
  }
  return NULL;
}

static void *ingest_thread_main(void *arg) {  // the queue side of proc_samples (radio.c:51-58)
  ka9q_rx *rx = arg;
  while (rx->running) {
    pthread_mutex_lock(&rx->qmutex);
    while (rx->queue == NULL && rx->running) {
      struct timespec ts;
      clock_gettime(CLOCK_REALTIME, &ts);
      ts.tv_nsec += 50000000L;
      if (ts.tv_nsec >= 1000000000L) {
        ts.tv_sec++;
        ts.tv_nsec -= 1000000000L;
      }
      pthread_cond_timedwait(&rx->qcond, &rx->qmutex, &ts);
    }
    pthread_mutex_unlock(&rx->qmutex);
    if (rx->running) ka9q_rx_drain(rx);
  }
  return NULL;
}

// fd: a bound (and multicast-joined) UDP socket owned by the caller; a receive timeout on it lets ka9q_rx_stop return fast
int ka9q_rx_start(ka9q_rx *rx, int fd) {
  if (!rx || fd < 0 || rx->threads_started) return -1;
  rx->fd = fd;
  rx->running = 1;
  struct timeval tv = {0, 100000};
  setsockopt(fd, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv));
  if (pthread_create(&rx->rx_thread, NULL, rx_thread_main, rx)) return -1;
  if (pthread_create(&rx->in_thread, NULL, ingest_thread_main, rx)) {
    rx->running = 0;
    pthread_join(rx->rx_thread, NULL);
    return -1;
  }
  rx->threads_started = 1;
  return 0;
}

int ka9q_rx_stop(ka9q_rx *rx) {
  if (!rx || !rx->threads_started) return 0;
  rx->running = 0;
  pthread_mutex_lock(&rx->qmutex);
  pthread_cond_broadcast(&rx->qcond);
  pthread_mutex_unlock(&rx->qmutex);
  pthread_mutex_lock(&rx->rmutex);
  pthread_cond_broadcast(&rx->rcond);
  pthread_mutex_unlock(&rx->rmutex);
  pthread_join(rx->rx_thread, NULL);
  pthread_join(rx->in_thread, NULL);
  rx->threads_started = 0;
  return 0;
}

void ka9q_rx_destroy(ka9q_rx *rx) {
  if (!rx) return;
  ka9q_rx_stop(rx);
  for (struct rx_packet *p = rx->queue; p;) {
    struct rx_packet *n = p->next;
    free(p);
    p = n;
  }
  for (struct rx_packet *p = rx->free_list; p;) {
    struct rx_packet *n = p->next;
    free(p);
    p = n;
  }
  (rx->pinned ? ka9q_host_free : ka9q_free)(rx->ring);
  free(rx->scratch);
  pthread_mutex_destroy(&rx->qmutex);
  pthread_cond_destroy(&rx->qcond);
  pthread_mutex_destroy(&rx->rmutex);
  pthread_cond_destroy(&rx->rcond);
  free(rx);
}

// ---- egress: the PCM of many channels for one block, packetised as audio.c:32-132 does and sent with sendmmsg ----
struct mm_ctx {
  int fd;
  struct mmsghdr *msgs;
  struct iovec *iov;
  unsigned char *store;
  int n, cap, sent, failed;
};
static int mm_flush(struct mm_ctx *c) {
  int off = 0;
  while (off < c->n) {
    const int r = sendmmsg(c->fd, c->msgs + off, (unsigned)(c->n - off), 0);
    if (r < 0) {
      if (errno == EINTR) continue;
      c->failed += c->n - off;  // e.g. ECONNREFUSED on a unicast socket without a listener (audio.c:73-77 gives up too)
      break;
    }
    c->sent += r;
    off += r;
  }
  c->n = 0;
  return 0;
}
static int mm_emit(void *user, const void *packet, int len) {
  struct mm_ctx *c = user;
  unsigned char *dst = c->store + (size_t)c->n * (12 + 2 * KA9Q_PCM_BUFSIZE);
  memcpy(dst, packet, (size_t)len);
  c->iov[c->n].iov_base = dst;
  c->iov[c->n].iov_len = (size_t)len;
  memset(&c->msgs[c->n], 0, sizeof(c->msgs[c->n]));
  c->msgs[c->n].msg_hdr.msg_iov = &c->iov[c->n];
  c->msgs[c->n].msg_hdr.msg_iovlen = 1;
  if (++c->n == c->cap) mm_flush(c);
  return 0;
}

// pcm_row: one block row of ka9q_stream_fetch; channel c's samples start at offs[c] (ka9q_stream_pcm_offset), `frames`
// frames of channels[c] interleaved int16. fd: a connected UDP socket. Returns packets sent, -1 on bad arguments.
int ka9q_pcm_send_block(int fd, ka9q_pcm_out *outs, const int16_t *pcm_row, const int *offs, const int *channels, int nchan,
                        int frames, int batch) {
  if (fd < 0 || !outs || !pcm_row || !offs || !channels || nchan < 1 || frames < 1) return -1;
  if (batch < 1) batch = 256;
  if (batch > 1024) batch = 1024;  // UIO_MAXIOV
  struct mm_ctx c;
  memset(&c, 0, sizeof(c));
  c.fd = fd;
  c.cap = batch;
  c.msgs = malloc(sizeof(*c.msgs) * (size_t)batch);
  c.iov = malloc(sizeof(*c.iov) * (size_t)batch);
  c.store = malloc((size_t)batch * (12 + 2 * KA9Q_PCM_BUFSIZE));
  if (!c.msgs || !c.iov || !c.store) {
    free(c.msgs);
    free(c.iov);
    free(c.store);
    return -1;
  }
  for (int ch = 0; ch < nchan; ch++) ka9q_pcm_packetise(&outs[ch], pcm_row + offs[ch], frames, channels[ch], mm_emit, &c);
  mm_flush(&c);
  free(c.msgs);
  free(c.iov);
  free(c.store);
  return c.sent;
}
