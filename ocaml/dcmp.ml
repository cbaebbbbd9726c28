(* ocaml/dcmp.ml — the externals of dcmp_stubs.c, shared by de.ml / zl.ml / gz.ml / lzo.ml. *)
type bigstring = (char, Bigarray.int8_unsigned_elt, Bigarray.c_layout) Bigarray.Array1.t
type i64s = (int64, Bigarray.int64_elt, Bigarray.c_layout) Bigarray.Array1.t
type i32s = (int32, Bigarray.int32_elt, Bigarray.c_layout) Bigarray.Array1.t

let fmt_raw = 0 and fmt_zlib = 1 and fmt_gzip = 2   (* DCMP_FMT_* *)

external pin : bigstring -> unit = "dcmp_stub_pin"
external unpin : bigstring -> unit = "dcmp_stub_unpin"
external inf_ns_inflate : int -> bigstring -> bigstring -> int * int * int = "dcmp_stub_inf_ns_inflate"
external inflate_batch : int -> bigstring -> i64s -> i64s -> bigstring -> i64s -> i64s -> i64s -> i64s -> i32s -> unit
  = "dcmp_stub_inflate_batch_bytecode" "dcmp_stub_inflate_batch"
external def_ns_deflate : int -> int -> bigstring -> bigstring -> int * int = "dcmp_stub_def_ns_deflate"
external lzo_uncompress : int -> bigstring -> bigstring -> int * int = "dcmp_stub_lzo_uncompress"

type inf
external inf_new : int -> bigstring -> inf = "dcmp_stub_inf_new"
external inf_src : inf -> bigstring -> int -> int -> unit = "dcmp_stub_inf_src"
external inf_decode : inf -> int = "dcmp_stub_inf_decode"
external inf_error : inf -> string = "dcmp_stub_inf_error"
external inf_dst_rem : inf -> int = "dcmp_stub_inf_dst_rem"
external inf_src_rem : inf -> int = "dcmp_stub_inf_src_rem"
external inf_write : inf -> int = "dcmp_stub_inf_write"
external inf_flush : inf -> unit = "dcmp_stub_inf_flush"
external inf_reset : inf -> unit = "dcmp_stub_inf_reset"
external inf_checksum : inf -> int32 = "dcmp_stub_inf_checksum"
external gz_inf_os : inf -> int = "dcmp_stub_gz_inf_os"
external gz_inf_field : inf -> int -> string option = "dcmp_stub_gz_inf_field"

type def
external def_new : int -> int -> def = "dcmp_stub_def_new"
external def_gz_header : def -> int32 -> int -> int -> string option -> string option -> unit
  = "dcmp_stub_def_gz_header_bytecode" "dcmp_stub_def_gz_header"
external def_src : def -> bigstring -> int -> int -> unit = "dcmp_stub_def_src"
external def_dst : def -> bigstring -> int -> int -> unit = "dcmp_stub_def_dst"
external def_encode : def -> int = "dcmp_stub_def_encode"
external def_src_rem : def -> int = "dcmp_stub_def_src_rem"
external def_dst_rem : def -> int = "dcmp_stub_def_dst_rem"

(* A streaming decoder as the shims use it: the handle plus what it points into (kept alive here), and the `String /
   `Channel sources of the reference driven on top of the `Manual protocol (lib/de.ml:744-753). *)
type src = [ `Channel of in_channel | `String of string | `Manual ]

type decoder = {
  h : inf;
  o : bigstring;
  source : src;
  mutable i : bigstring;           (* the input handed over last *)
  mutable string_fed : bool;
  ibuf : bigstring Lazy.t;         (* `Channel: io_buffer_size bytes *)
}

let bigstring_of_string s =
  let b = Bigarray.Array1.create Bigarray.char Bigarray.c_layout (String.length s) in
  String.iteri (fun k c -> Bigarray.Array1.unsafe_set b k c) s ; b

let empty = Bigarray.Array1.create Bigarray.char Bigarray.c_layout 0

let decoder fmt (source : src) ~o =
  { h = inf_new fmt o; o; source; i = empty; string_fed = false;
    ibuf = lazy (Bigarray.Array1.create Bigarray.char Bigarray.c_layout 65536) }

let src d s j l = d.i <- s ; inf_src d.h s j l

(* one step of the protocol: `Await is answered here for `String and `Channel sources *)
let rec decode d : [ `Await | `Flush | `End | `Malformed of string ] =
  match inf_decode d.h with
  | 1 -> `Flush
  | 2 -> `End
  | 3 -> `Malformed (inf_error d.h)
  | _ -> (
    match d.source with
    | `Manual -> `Await
    | `String s ->
      if d.string_fed then src d empty 0 0
      else (d.string_fed <- true ; let b = bigstring_of_string s in src d b 0 (Bigarray.Array1.dim b))
      ; decode d
    | `Channel ic ->
      let buf = Lazy.force d.ibuf in
      let tmp = Bytes.create (Bigarray.Array1.dim buf) in
      let n = input ic tmp 0 (Bytes.length tmp) in
      for k = 0 to n - 1 do Bigarray.Array1.unsafe_set buf k (Bytes.unsafe_get tmp k) done
      ; src d buf 0 n ; decode d)

let reset d = inf_reset d.h ; d.string_fed <- false
