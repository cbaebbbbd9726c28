/* oracle/lzo.c — restatement of the LZO1X decoder of lib/lzo.ml.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).
 *
 * Follows /root/reference/lib/lzo.ml:
 *   transmit / copy            lzo.ml:188-197
 *   count                      lzo.ml:218-234
 *   copy_done                  lzo.ml:242-244
 *   run (the interpreter)      lzo.ml:246-293
 *   fiber (main opcode loop)   lzo.ml:315-368
 *   fiber (first byte)         lzo.ml:372-393
 *   uncompress                 lzo.ml:395-403
 *   uncompress_with_buffer     lzo.ml:405-414
 *
 * Deliberate choice: back-copies are byte-replicating,
 * i.e. what `uncompress_with_buffer` (lzo.ml:206-216), minilzo and every test
 * pin; the bigstring `uncompress` copies overlapping 4-byte words and leaks
 * stale output bytes for off < 4 (lzo.ml:65-79).
 *
 * mode 0 = error mapping of `uncompress`           (Out_of_bound -> Invalid_argument)
 * mode 1 = error mapping of `uncompress_with_buffer` (Out_of_bound -> Malformed,
 *          distance beyond produced output -> Invalid_dictionary, no output cap)
 */
#include "oracle.h"

typedef struct {
  const uint8_t *i; int64_t i_pos, i_len;
  uint8_t *o; int64_t o_pos, o_len;
  int64_t state;
  int mode;
} lz_t;

#define E_OK 0
#define E_MALFORMED ORC_LZO_MALFORMED
#define E_OOB (-1) /* exception Out_of_bound, mapped at the end */
#define E_DICT ORC_LZO_INVALID_DICTIONARY

/* lzo.ml:188-192 (+ blit bounds check lzo.ml:81-90) */
static int transmit(lz_t *v, int64_t len) {
  int64_t k;
  if (len < 0 || v->i_pos > v->i_len - len) return E_OOB;
  if (v->o_pos > v->o_len - len) return E_OOB;
  for (k = 0; k < len; k++) v->o[v->o_pos + k] = v->i[v->i_pos + k];
  v->i_pos += len; v->o_pos += len;
  return E_OK;
}

/* lzo.ml:194-197 / 206-216 */
static int copy(lz_t *v, int64_t off, int64_t len) {
  int64_t k, s = v->o_pos - off;
  if (v->mode == 1) { if (v->o_pos < off) return E_DICT; }
  else if (s < 0) return E_OOB;
  if (v->o_pos > v->o_len - len) return E_OOB;
  for (k = 0; k < len; k++) v->o[v->o_pos + k] = v->o[s + k];
  v->o_pos += len;
  return E_OK;
}

/* lzo.ml:218-234 */
static int count(lz_t *t, int64_t *out) {
  int64_t res = 0, idx = t->i_pos, max = t->i_len;
  while (idx < max && t->i[idx] == 0) { idx++; res++; }
  if (idx < max) { t->i_pos = idx + 1; *out = res * 255 + t->i[idx]; return E_OK; }
  return E_MALFORMED;
}

#define NEED_BYTE(t) do { if ((t)->i_pos >= (t)->i_len) return E_MALFORMED; } while (0)
/* Peek/Junk (Short _): guard lzo.ml:267-268, then get_int16 bounds lzo.ml:133-135 */
#define NEED_SHORT(t) do { if ((t)->i_pos >= (t)->i_len) return E_MALFORMED; \
                           if ((t)->i_pos > (t)->i_len - 2) return E_OOB; } while (0)

/* Copy (sub, state) in run, lzo.ml:286-290: state set, len + 2 bytes, then state&3 literals */
static int do_copy(lz_t *t, int64_t off, int64_t len, int64_t state) {
  int r;
  if (t->i_pos >= t->i_len) return E_MALFORMED; /* guard of `go`, lzo.ml:267-268 */
  t->state = state;
  r = copy(t, off, len + 2);
  if (r) return r;
  return transmit(t, t->state & 3);
}

/* lzo.ml:315-368 */
static int main_loop(lz_t *t) {
  for (;;) {
    int chr, st, r; int64_t cnt;
    NEED_BYTE(t);
    chr = t->i[t->i_pos++];
    st = (int)(t->state & 3); /* -1 land 3 = 3 */
    if (chr <= 15 && st == 0) {
      int64_t len;
      if (chr == 0) {
        NEED_BYTE(t);
        r = count(t, &cnt); if (r) return r;
        len = 3 + 15 + cnt;
      } else len = chr + 3;
      if (t->i_pos >= t->i_len) return E_MALFORMED; /* guard of `go`, lzo.ml:267-268 */
      t->state = -1;
      r = transmit(t, len); if (r) return r;
    } else if (chr <= 15) {
      /* st in {1,2,3}; the `-1` arm of lzo.ml:332-336 is unreachable */
      int d = chr >> 2, h;
      NEED_BYTE(t);
      h = t->i[t->i_pos++];
      r = do_copy(t, ((int64_t)h << 2) + d + 1, 0, chr & 3); if (r) return r;
    } else if (chr <= 31) {
      int64_t len = chr & 7, s, off;
      if (len == 0) { NEED_BYTE(t); r = count(t, &cnt); if (r) return r; len = 7 + cnt; }
      NEED_SHORT(t);
      s = t->i[t->i_pos] | (t->i[t->i_pos + 1] << 8); t->i_pos += 2;
      off = 16384 + ((int64_t)((chr & 8) >> 3) << 14) + (s >> 2);
      if (off == 16384) return E_OK;
      r = do_copy(t, off, len, s & 0xff); if (r) return r;
    } else if (chr <= 63) {
      int64_t len = chr & 31, s;
      if (len == 0) { NEED_BYTE(t); r = count(t, &cnt); if (r) return r; len = 31 + cnt; }
      NEED_SHORT(t);
      s = t->i[t->i_pos] | (t->i[t->i_pos + 1] << 8); t->i_pos += 2;
      r = do_copy(t, (s >> 2) + 1, len, s & 0xff); if (r) return r;
    } else {
      int len = (chr >> 5) - 1, d = (chr >> 2) & 7, h;
      NEED_BYTE(t);
      h = t->i[t->i_pos++];
      r = do_copy(t, ((int64_t)h << 3) + d + 1, len, chr); if (r) return r;
    }
  }
}

/* lzo.ml:372-393 */
static int first(lz_t *t) {
  int chr, r;
  NEED_BYTE(t);
  chr = t->i[t->i_pos];
  if (chr == 16) return E_MALFORMED; /* "No dictionary at offset 0 available" */
  if (chr <= 17) return main_loop(t);
  t->i_pos++;
  if (t->i_pos >= t->i_len) return E_MALFORMED; /* guard of `go`, lzo.ml:267-268 */
  t->state = 0;
  r = transmit(t, chr - 17); /* 18..21 -> 1..4, 22..255 -> chr - 17; all with State._0 */
  if (r) return r;
  return main_loop(t);
}

static int run(const uint8_t *src, int64_t src_len, uint8_t *dst, int64_t dst_len,
               int64_t *consumed, int64_t *produced, int mode) {
  lz_t v; int r;
  v.i = src; v.i_pos = 0; v.i_len = src_len;
  v.o = dst; v.o_pos = 0; v.o_len = dst_len;
  v.state = 0; v.mode = mode;
  r = first(&v);
  *consumed = v.i_pos; *produced = v.o_pos;
  if (r == E_OOB) return mode == 1 ? ORC_LZO_MALFORMED : ORC_LZO_INVALID_ARGUMENT;
  return r;
}

int orc_lzo_uncompress(const uint8_t *src, int64_t src_len, uint8_t *dst,
                       int64_t dst_len, int64_t *consumed, int64_t *produced) {
  return run(src, src_len, dst, dst_len, consumed, produced, 0);
}

int orc_lzo_uncompress_with_buffer(const uint8_t *src, int64_t src_len, uint8_t *dst,
                                   int64_t dst_len, int64_t *consumed, int64_t *produced) {
  return run(src, src_len, dst, dst_len, consumed, produced, 1);
}
