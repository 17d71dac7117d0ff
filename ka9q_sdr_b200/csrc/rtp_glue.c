// Wire-format glue either side of the DSP path: host C, no GPU, no sockets.
//   receive: main.c:313-344 (rtp_recv parsing), multicast.c:242-277 (ntoh_rtp), multicast.c:305-340 (rtp_process),
//            radio.c:60-100 (sample count, SSRC change, lost-sample zero fill)
//   send:    audio.c:32-132 (send_stereo_output / send_mono_output), multicast.c:282-294 (hton_rtp)
#include <stdint.h>
#include <string.h>
#include "../../include/ka9q_b200.h"

static inline uint16_t rd16(const unsigned char *p) { return (uint16_t)(p[0] << 8 | p[1]); }
static inline uint32_t rd32(const unsigned char *p) {
  return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | (uint32_t)p[3];
}

void ka9q_ingest_init(ka9q_ingest *g, int iq_format) {
  if (!g) return;
  memset(g, 0, sizeof(*g));
  g->iq_format = iq_format;
}

int ka9q_rtp_process(ka9q_rtp_state *state, uint32_t ssrc, uint16_t seq, uint32_t timestamp, int sampcnt) {
  if (ssrc != state->ssrc) {  // multicast.c:306-313: a new sender restarts the session
    state->init = 0;
    state->ssrc = ssrc;
  }
  if (!state->init) {  // multicast.c:314-321
    state->packets = 0;
    state->seq = seq;
    state->timestamp = timestamp;
    state->dupes = 0;
    state->drops = 0;
    state->init = 1;
  }
  state->packets++;
  const short seq_step = (short)(seq - state->seq);  // multicast.c:324-331
  if (seq_step != 0) {
    if (seq_step < 0) {
      state->dupes++;
      return -1;
    }
    state->drops += seq_step;
  }
  state->seq = (uint16_t)(seq + 1);
  const int time_step = (int)(timestamp - state->timestamp);  // multicast.c:334-339
  if (time_step < 0) return time_step;
  state->timestamp = timestamp + (uint32_t)sampcnt;
  return time_step;
}

long long ka9q_ingest_datagram(ka9q_ingest *g, const void *datagram, int size, void *dst, long long room) {
  if (!g || !datagram || !dst) return -1;
  if (size < KA9Q_RTP_MIN_SIZE) {  // main.c:319-320
    g->ignored++;
    return -1;
  }
  const unsigned char *dp = (const unsigned char *)datagram;
  const unsigned char *const end = dp + size;
  // ntoh_rtp (multicast.c:242-277)
  const int pad = (dp[0] >> 5) & 1, extension = (dp[0] >> 4) & 1, cc = dp[0] & 0xf;
  const int type = dp[1] & 0x7f;
  const uint16_t seq = rd16(dp + 2);
  const uint32_t timestamp = rd32(dp + 4), ssrc = rd32(dp + 8);
  dp += 12 + 4 * cc;
  if (extension) {
    if (dp + 4 > end) {
      g->ignored++;
      return -1;
    }
    dp += 4 + 4 + rd16(dp + 2);  // type, length, then "4 + length" bytes exactly as multicast.c:271-274 skips
  }
  if (dp > end) {  // the reference would read past the datagram here; a malformed header is simply ignored
    g->ignored++;
    return -1;
  }
  long long len = end - dp;
  if (pad && len > 0) len -= end[-1];  // main.c:326-330
  if (type != KA9Q_IQ_PT && type != KA9Q_IQ_PT8) {  // main.c:331-332
    g->ignored++;
    return -1;
  }
  dp += 24;  // legacy status header, host byte order, ignored (main.c:340-341)
  len -= 24;
  const int bytes_per_sample = (type == KA9Q_IQ_PT) ? 4 : 2;
  const int want = (g->iq_format == KA9Q_IQ_S16) ? 4 : 2;
  if (bytes_per_sample != want || len < 0) {  // one sample format per stream (the reference converts per packet)
    g->ignored++;
    return -1;
  }
  const int sampcount = (int)(len / bytes_per_sample);  // radio.c:62-71
  // room check before the RTP state moves, on a copy of the state
  ka9q_rtp_state st = g->rtp;
  const int ssrc_changed = ssrc != st.ssrc;
  const int time_step = ka9q_rtp_process(&st, ssrc, seq, timestamp, sampcount);  // radio.c:77
  if (time_step < 0 || time_step > 192000) {  // radio.c:78-81: old samples or too big a jump: drop
    g->rtp = st;
    if (ssrc_changed) g->samples = 0;
    g->ignored++;
    return -1;
  }
  const long long total = (long long)time_step + sampcount;
  if (total > room) return -2;
  g->rtp = st;
  if (ssrc_changed) g->samples = 0;  // radio.c:72-76
  unsigned char *out = (unsigned char *)dst;
  if (time_step > 0) {  // radio.c:82-100: zeros keep the sample count (and with it every LO phase) right
    memset(out, 0, (size_t)time_step * want);
    out += (size_t)time_step * want;
    g->zero_filled += time_step;
  }
  memcpy(out, dp, (size_t)sampcount * want);
  g->samples += total;
  return total;
}

int ka9q_pcm_packetise(ka9q_pcm_out *o, const int16_t *pcm, int frames, int channels, ka9q_emit_fn emit, void *user) {
  if (!o || !pcm || !emit || frames < 0 || (channels != 1 && channels != 2)) return -1;
  int sent = 0;
  int size = frames;
  while (size > 0) {
    // audio.c:46 / :95: chunk counts int16 words; stereo packs 240 frames, mono 480
    const int chunk = channels == 2 ? (2 * size < KA9Q_PCM_BUFSIZE ? 2 * size : KA9Q_PCM_BUFSIZE)
                                    : (size < KA9Q_PCM_BUFSIZE ? size : KA9Q_PCM_BUFSIZE);
    unsigned char packet[12 + 2 * KA9Q_PCM_BUFSIZE];
    unsigned char *dp = packet + 12;
    int not_silent = 0;
    for (int i = 0; i < chunk; i++) {
      const uint16_t w = (uint16_t)pcm[i];
      not_silent |= w;
      *dp++ = (unsigned char)(w >> 8);  // htons
      *dp++ = (unsigned char)w;
    }
    pcm += chunk;
    const uint32_t ts = o->rtp.timestamp;
    o->rtp.timestamp += (uint32_t)(chunk / channels);  // frames, also for suppressed chunks (audio.c:54-56,:104-106)
    if (not_silent) {
      o->rtp.packets++;
      o->rtp.bytes += 2 * chunk;
      const int marker = o->silent ? 1 : 0;  // first packet after silence (audio.c:59-63)
      o->silent = 0;
      const uint16_t seq = o->rtp.seq++;
      // hton_rtp (multicast.c:282-294): version 2, no padding / extension / CSRC
      packet[0] = 2 << 6;
      packet[1] = (unsigned char)((marker << 7) | (channels == 2 ? KA9Q_PCM_STEREO_PT : KA9Q_PCM_MONO_PT));
      packet[2] = (unsigned char)(seq >> 8);
      packet[3] = (unsigned char)seq;
      packet[4] = (unsigned char)(ts >> 24);
      packet[5] = (unsigned char)(ts >> 16);
      packet[6] = (unsigned char)(ts >> 8);
      packet[7] = (unsigned char)ts;
      packet[8] = (unsigned char)(o->rtp.ssrc >> 24);
      packet[9] = (unsigned char)(o->rtp.ssrc >> 16);
      packet[10] = (unsigned char)(o->rtp.ssrc >> 8);
      packet[11] = (unsigned char)o->rtp.ssrc;
      if (emit(user, packet, 12 + 2 * chunk) < 0) break;  // audio.c:73-77: give up on a send error
      sent++;
    } else {
      o->silent = 1;
    }
    size -= chunk / channels;
  }
  return sent;
}

// ---- status TLV (status.c:31-96, radio_status.c:171-203; enum status_type status.h:6-70) ----
enum {
  // positions in enum status_type (status.h:6-70); tests/test_rtp_glue.py re-derives them from the reference header
  ST_EOL = 0, ST_NOISE_BANDWIDTH = 35, ST_IF_POWER = 36, ST_BASEBAND_POWER = 37, ST_DEMOD_MODE = 40,
  ST_INDEPENDENT_SIDEBAND = 41, ST_DEMOD_SNR = 42, ST_DEMOD_GAIN = 43, ST_FREQ_OFFSET = 44, ST_PEAK_DEVIATION = 45,
  ST_OUTPUT_CHANNELS = 50
};

// encode_int64 (status.c:31-51): type, length, value big-endian with the leading zero bytes dropped
static unsigned char *tlv_int(unsigned char *cp, int type, uint64_t x) {
  *cp++ = (unsigned char)type;
  int len = 8;
  while (len > 0 && (x & 0xff00000000000000ULL) == 0) {
    x <<= 8;
    len--;
  }
  *cp++ = (unsigned char)len;
  for (int i = 0; i < len; i++) {
    *cp++ = (unsigned char)(x >> 56);
    x <<= 8;
  }
  return cp;
}
static unsigned char *tlv_float(unsigned char *cp, int type, float x) {  // encode_float (status.c:83-88)
  uint32_t d;
  memcpy(&d, &x, sizeof d);
  return tlv_int(cp, type, d);
}
static unsigned char *tlv_byte(unsigned char *cp, int type, unsigned char x) {  // encode_byte (status.c:62-69)
  *cp++ = (unsigned char)type;
  *cp++ = 1;
  *cp++ = x;
  return cp;
}

int ka9q_status_encode_signals(const ka9q_chan_status *st, int demod_type, int isb, float if_power, float noise_bandwidth,
                               int output_channels, unsigned char *buf, int room) {
  if (!st || !buf) return -1;
  unsigned char tmp[96], *cp = tmp;  // at most 10 items of <= 6 bytes and the EOL
  cp = tlv_float(cp, ST_NOISE_BANDWIDTH, noise_bandwidth);    // radio_status.c:171
  cp = tlv_float(cp, ST_IF_POWER, if_power);                  // :174
  cp = tlv_float(cp, ST_BASEBAND_POWER, st->bb_power);        // :175
  cp = tlv_byte(cp, ST_DEMOD_MODE, (unsigned char)demod_type);  // :181
  switch (demod_type) {
    case 1:  // AM_DEMOD (:183-185)
      cp = tlv_float(cp, ST_DEMOD_GAIN, st->agc_gain);
      break;
    case 2:  // FM_DEMOD (:186-191; PL_TONE is not computed here)
      cp = tlv_float(cp, ST_PEAK_DEVIATION, st->pdeviation);
      cp = tlv_float(cp, ST_FREQ_OFFSET, st->foffset);
      cp = tlv_float(cp, ST_DEMOD_SNR, st->snr);
      break;
    default:  // LINEAR_DEMOD (:192-202, no PLL)
      cp = tlv_float(cp, ST_DEMOD_GAIN, st->agc_gain);
      cp = tlv_int(cp, ST_INDEPENDENT_SIDEBAND, (uint32_t)isb);
      break;
  }
  cp = tlv_int(cp, ST_OUTPUT_CHANNELS, (uint32_t)output_channels);  // :204
  *cp++ = ST_EOL;                                                   // :205
  const int n = (int)(cp - tmp);
  if (n > room) return -1;
  memcpy(buf, tmp, (size_t)n);
  return n;
}
