/* ocaml/dcmp_stubs.c — C stubs between the OCaml shims (ocaml/{de,zl,gz,lzo}.ml) and libdcmp_b200.so (include/dcmp.h).
 *
 * mirage/decompress has no FFI (pure OCaml); this file IS the binding a maintainer adds to put the B200 engine under
 * lib/de.mli, lib/zl.mli, lib/gz.mli, lib/lzo.mli.  `bigstring` = (char, int8_unsigned_elt, c_layout) Bigarray.Array1.t
 * (lib/de.mli:45) is stable malloc'd memory: Caml_ba_data_val crosses the boundary without a copy.  Every call that
 * launches kernels releases the OCaml runtime lock.  The OCaml toolchain is not in the build image: the file is
 * source for the reference's build (dune: foreign_stubs), not compiled here; the same C entry points are exercised
 * through ctypes by decompress_b200/*.py in tests/. */
#include <string.h>
#include <caml/mlvalues.h>
#include <caml/memory.h>
#include <caml/alloc.h>
#include <caml/bigarray.h>
#include <caml/custom.h>
#include <caml/fail.h>
#include <caml/threads.h>
#include "dcmp.h"

static dcmp_engine *engine(void) { /* one engine (CUDA device, streams, scratch) per process */
  static dcmp_engine *e = NULL;
  if (!e && dcmp_engine_create(0, &e) != DCMP_API_OK) caml_failwith(dcmp_last_error());
  return e;
}
static void api_check(int rc) { if (rc < 0) caml_failwith(dcmp_last_error()); }
#define BA_PTR(v) ((uint8_t *)Caml_ba_data_val(v))
#define BA_LEN(v) ((int64_t)Caml_ba_array_val(v)->dim[0])

/* Dcmp.pin / unpin : bigstring -> unit — page-lock a long-lived io buffer (a Bigarray is pageable malloc memory; the
 * *_batch calls copy from / to it at the PCIe rate only when it is registered) */
CAMLprim value dcmp_stub_pin(value ba) { api_check(dcmp_host_register(BA_PTR(ba), BA_LEN(ba))); return Val_unit; }
CAMLprim value dcmp_stub_unpin(value ba) { api_check(dcmp_host_unregister(BA_PTR(ba))); return Val_unit; }

/* ---- one-shot: De.Inf.Ns.inflate / Zl.Inf.Ns.inflate (lib/de.ml:1807-1822, lib/zl.ml:400-417) ----------------------
 * fmt -> src -> dst -> (status, consumed, produced) */
CAMLprim value dcmp_stub_inf_ns_inflate(value fmt, value src, value dst) {
  CAMLparam3(fmt, src, dst);
  CAMLlocal1(res);
  const uint8_t *s = BA_PTR(src);
  uint8_t *d = BA_PTR(dst);
  int64_t sl = BA_LEN(src), dl = BA_LEN(dst), zero = 0, cons = 0, prod = 0;
  int32_t st = 0;
  const int f = Int_val(fmt);
  dcmp_engine *e = engine();
  caml_release_runtime_system();
  int rc = dcmp_inflate_batch(e, f, 1, s, &zero, &sl, d, &zero, &dl, &cons, &prod, &st);
  caml_acquire_runtime_system();
  api_check(rc);
  res = caml_alloc_tuple(3);
  Store_field(res, 0, Val_int(st));
  Store_field(res, 1, Val_long(cons));
  Store_field(res, 2, Val_long(prod));
  CAMLreturn(res);
}

/* batched form: slabs + int64 Bigarray tables (the shape the GPU wants): n streams, one call */
CAMLprim value dcmp_stub_inflate_batch(value fmt, value src, value soff, value slen, value dst, value doff, value dcap,
                                       value cons, value prod, value st) {
  CAMLparam5(fmt, src, soff, slen, dst);
  CAMLxparam5(doff, dcap, cons, prod, st);
  const int64_t n = BA_LEN(soff);
  const int f = Int_val(fmt);
  void *p[9] = {BA_PTR(src), BA_PTR(soff), BA_PTR(slen), BA_PTR(dst), BA_PTR(doff), BA_PTR(dcap), BA_PTR(cons), BA_PTR(prod), BA_PTR(st)};
  dcmp_engine *e = engine();
  caml_release_runtime_system();
  int rc = dcmp_inflate_batch(e, f, n, p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], p[8]);
  caml_acquire_runtime_system();
  api_check(rc);
  CAMLreturn(Val_unit);
}
CAMLprim value dcmp_stub_inflate_batch_bytecode(value *a, int n) {
  (void)n;
  return dcmp_stub_inflate_batch(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9]);
}

/* De.Def.Ns.deflate / Zl.Def.Ns.deflate (lib/de.ml:3998-4010, lib/zl.ml:610-630): fmt -> level -> src -> dst -> (status, len) */
CAMLprim value dcmp_stub_def_ns_deflate(value fmt, value level, value src, value dst) {
  CAMLparam4(fmt, level, src, dst);
  CAMLlocal1(res);
  const uint8_t *s = BA_PTR(src);
  uint8_t *d = BA_PTR(dst);
  int64_t sl = BA_LEN(src), dl = BA_LEN(dst), written = 0;
  const int f = Int_val(fmt), lv = Int_val(level);
  dcmp_engine *e = engine();
  caml_release_runtime_system();
  int st = f == DCMP_FMT_ZLIB ? dcmp_zl_def_ns_deflate(e, lv, s, sl, d, dl, &written) : dcmp_def_ns_deflate(e, lv, s, sl, d, dl, &written);
  caml_acquire_runtime_system();
  api_check(st);
  res = caml_alloc_tuple(2);
  Store_field(res, 0, Val_int(st));
  Store_field(res, 1, Val_long(written));
  CAMLreturn(res);
}

/* Lzo.uncompress (lib/lzo.ml:395-403): mode 0; uncompress_with_buffer (405-414): mode 1.  -> (status, produced) */
CAMLprim value dcmp_stub_lzo_uncompress(value mode, value src, value dst) {
  CAMLparam3(mode, src, dst);
  CAMLlocal1(res);
  const uint8_t *s = BA_PTR(src);
  uint8_t *d = BA_PTR(dst);
  int64_t sl = BA_LEN(src), dl = BA_LEN(dst), zero = 0, cons = 0, prod = 0;
  int32_t st = 0;
  const int m = Int_val(mode);
  dcmp_engine *e = engine();
  caml_release_runtime_system();
  int rc = dcmp_lzo_uncompress_batch(e, m, 1, s, &zero, &sl, d, &zero, &dl, &cons, &prod, &st);
  caml_acquire_runtime_system();
  api_check(rc);
  res = caml_alloc_tuple(2);
  Store_field(res, 0, Val_int(st));
  Store_field(res, 1, Val_long(prod));
  CAMLreturn(res);
}

/* ---- streaming decoders: De.Inf.decoder / Zl.Inf / Gz.Inf over dcmp_inf_* ---------------------------------------------
 * The handle is a custom block; the bigstrings it points into (`o`, the current input) are kept alive by the OCaml
 * record that holds the handle (ocaml/de.ml). */
#define Inf_val(v) (*((dcmp_inf **)Data_custom_val(v)))
static void inf_finalize(value v) { dcmp_inf_free(Inf_val(v)); }
static struct custom_operations inf_ops = {"dcmp.inf", inf_finalize, custom_compare_default, custom_hash_default,
                                           custom_serialize_default, custom_deserialize_default, custom_compare_ext_default,
                                           custom_fixed_length_default};
CAMLprim value dcmp_stub_inf_new(value fmt, value o) {
  CAMLparam2(fmt, o);
  CAMLlocal1(v);
  dcmp_inf *d = NULL;
  api_check(dcmp_inf_new(engine(), Int_val(fmt), BA_PTR(o), BA_LEN(o), &d));
  v = caml_alloc_custom(&inf_ops, sizeof(dcmp_inf *), 0, 1);
  Inf_val(v) = d;
  CAMLreturn(v);
}
CAMLprim value dcmp_stub_inf_src(value h, value buf, value off, value len) { /* raises Invalid_argument like de.ml:736-737 */
  if (Long_val(off) < 0 || Long_val(len) < 0 || Long_val(off) + Long_val(len) > BA_LEN(buf)) caml_invalid_argument("Invalid bounds");
  api_check(dcmp_inf_src(Inf_val(h), BA_PTR(buf) + Long_val(off), Long_val(len)));
  return Val_unit;
}
CAMLprim value dcmp_stub_inf_decode(value h) { /* 0 `Await, 1 `Flush, 2 `End, 3 `Malformed */
  CAMLparam1(h);
  dcmp_inf *d = Inf_val(h);
  caml_release_runtime_system();
  int r = dcmp_inf_decode(d);
  caml_acquire_runtime_system();
  api_check(r);
  CAMLreturn(Val_int(r));
}
CAMLprim value dcmp_stub_inf_error(value h) { return caml_copy_string(dcmp_inf_error(Inf_val(h))); }
CAMLprim value dcmp_stub_inf_dst_rem(value h) { return Val_long(dcmp_inf_dst_rem(Inf_val(h))); }
CAMLprim value dcmp_stub_inf_src_rem(value h) { return Val_long(dcmp_inf_src_rem(Inf_val(h))); }
CAMLprim value dcmp_stub_inf_write(value h) { return Val_long(dcmp_inf_write(Inf_val(h))); }
CAMLprim value dcmp_stub_inf_flush(value h) { dcmp_inf_flush(Inf_val(h)); return Val_unit; }
CAMLprim value dcmp_stub_inf_reset(value h) { dcmp_inf_reset(Inf_val(h)); return Val_unit; }
CAMLprim value dcmp_stub_inf_checksum(value h) { return caml_copy_int32((int32_t)dcmp_inf_checksum(Inf_val(h))); }
CAMLprim value dcmp_stub_gz_inf_os(value h) { return Val_int(dcmp_gz_inf_os(Inf_val(h))); }
CAMLprim value dcmp_stub_gz_inf_field(value h, value which) { /* string option */
  CAMLparam2(h, which);
  CAMLlocal2(s, some);
  const uint8_t *p = NULL;
  int64_t n = dcmp_gz_inf_field(Inf_val(h), Int_val(which), &p);
  if (n < 0) CAMLreturn(Val_int(0));
  s = caml_alloc_initialized_string((mlsize_t)n, (const char *)p);
  some = caml_alloc(1, 0);
  Store_field(some, 0, s);
  CAMLreturn(some);
}

/* ---- streaming encoders: Zl.Def / Gz.Def / De.Higher.compress over dcmp_def_* -------------------------------------- */
#define Def_val(v) (*((dcmp_def **)Data_custom_val(v)))
static void def_finalize(value v) { dcmp_def_free(Def_val(v)); }
static struct custom_operations def_ops = {"dcmp.def", def_finalize, custom_compare_default, custom_hash_default,
                                           custom_serialize_default, custom_deserialize_default, custom_compare_ext_default,
                                           custom_fixed_length_default};
CAMLprim value dcmp_stub_def_new(value fmt, value level) {
  CAMLparam2(fmt, level);
  CAMLlocal1(v);
  dcmp_def *d = NULL;
  int rc = dcmp_def_new(engine(), Int_val(fmt), Int_val(level), &d);
  if (rc == DCMP_API_INVALID_ARGUMENT) caml_invalid_argument("Invalid level"); /* De.Lz77.state, de.ml:4431-4437 */
  api_check(rc);
  v = caml_alloc_custom(&def_ops, sizeof(dcmp_def *), 0, 1);
  Def_val(v) = d;
  CAMLreturn(v);
}
/* mtime -> os -> flags (1 ascii, 2 hcrc) -> filename option -> comment option */
CAMLprim value dcmp_stub_def_gz_header(value h, value mtime, value os, value flags, value fname, value fcomment) {
  const uint8_t *fn = Is_block(fname) ? (const uint8_t *)String_val(Field(fname, 0)) : NULL;
  const uint8_t *cm = Is_block(fcomment) ? (const uint8_t *)String_val(Field(fcomment, 0)) : NULL;
  api_check(dcmp_def_gz_header(Def_val(h), (uint32_t)Int32_val(mtime), Int_val(os), Int_val(flags) & 1, (Int_val(flags) >> 1) & 1, fn,
                               fn ? (int64_t)caml_string_length(Field(fname, 0)) : 0, cm, cm ? (int64_t)caml_string_length(Field(fcomment, 0)) : 0));
  return Val_unit;
}
CAMLprim value dcmp_stub_def_gz_header_bytecode(value *a, int n) { (void)n; return dcmp_stub_def_gz_header(a[0], a[1], a[2], a[3], a[4], a[5]); }
CAMLprim value dcmp_stub_def_src(value h, value buf, value off, value len) {
  if (Long_val(off) < 0 || Long_val(len) < 0 || Long_val(off) + Long_val(len) > BA_LEN(buf)) caml_invalid_argument("Invalid bounds");
  api_check(dcmp_def_src(Def_val(h), BA_PTR(buf) + Long_val(off), Long_val(len)));
  return Val_unit;
}
CAMLprim value dcmp_stub_def_dst(value h, value buf, value off, value len) {
  if (Long_val(off) < 0 || Long_val(len) < 0 || Long_val(off) + Long_val(len) > BA_LEN(buf)) caml_invalid_argument("Invalid bounds");
  api_check(dcmp_def_dst(Def_val(h), BA_PTR(buf) + Long_val(off), Long_val(len)));
  return Val_unit;
}
CAMLprim value dcmp_stub_def_encode(value h) { /* 0 `Await, 1 `Flush, 2 `End */
  CAMLparam1(h);
  dcmp_def *d = Def_val(h);
  caml_release_runtime_system();
  int r = dcmp_def_encode(d);
  caml_acquire_runtime_system();
  api_check(r);
  CAMLreturn(Val_int(r));
}
CAMLprim value dcmp_stub_def_src_rem(value h) { return Val_long(dcmp_def_src_rem(Def_val(h))); }
CAMLprim value dcmp_stub_def_dst_rem(value h) { return Val_long(dcmp_def_dst_rem(Def_val(h))); }
