(* ocaml/zl.ml — drop-in body for lib/zl.ml behind the UNCHANGED lib/zl.mli: Zl.Inf / Zl.Def / Zl.Higher over
 * libdcmp_b200.so (dcmp_inf_* / dcmp_def_* with DCMP_FMT_ZLIB; the zlib header and the Adler-32 trailer are handled
 * on the host side of the C library, the DEFLATE data on the B200).  The reference threads a persistent value through
 * decode / src / flush (zl.mli:39-108); here the value is a handle on mutable C state and every function returns the
 * decoder it was given: callers that use each version once — every caller in the reference's tree does — see no
 * difference. *)
type bigstring = De.bigstring
type window = De.window

let io_buffer_size = De.io_buffer_size

module Inf = struct
  type src = Dcmp.src
  type decoder = Dcmp.decoder
  type signal = [ `Await of decoder | `Flush of decoder | `End of decoder | `Malformed of string ]

  let decoder (src : src) ~o ~allocate:(_ : int -> window) : decoder = Dcmp.decoder Dcmp.fmt_zlib src ~o

  let decode (d : decoder) : signal =               (* was lib/zl.ml:282-309 *)
    match Dcmp.decode d with
    | `Await -> `Await d | `Flush -> `Flush d | `End -> `End d | `Malformed err -> `Malformed err

  let reset d = Dcmp.reset d ; d
  let src d s j l = Dcmp.src d s j l ; d
  let flush (d : decoder) = Dcmp.inf_flush d.Dcmp.h ; d
  let dst_rem (d : decoder) = Dcmp.inf_dst_rem d.Dcmp.h
  let src_rem (d : decoder) = Dcmp.inf_src_rem d.Dcmp.h
  let write (d : decoder) = Dcmp.inf_write d.Dcmp.h

  module Ns = struct
    type error = [ `Invalid_header | `Invalid_checksum | De.Inf.Ns.error ]

    let pp_error ppf = function
      | `Invalid_header -> Format.pp_print_string ppf "Invalid Zlib header"
      | `Invalid_checksum -> Format.pp_print_string ppf "Invalid checksum"
      | #De.Inf.Ns.error as e -> De.Inf.Ns.pp_error ppf e

    let inflate src dst =                           (* was lib/zl.ml:400-417 *)
      match Dcmp.inf_ns_inflate Dcmp.fmt_zlib src dst with
      | 0, consumed, produced -> Ok (consumed, produced)
      | 9, _, _ -> Error `Invalid_header
      | 10, _, _ -> Error `Invalid_checksum
      | st, _, _ -> Error (De.Inf.Ns.error_of_status st :> [> error ])
  end
end

module Def = struct
  type src = [ `Channel of in_channel | `String of string | `Manual ]
  type dst = [ `Channel of out_channel | `Buffer of Buffer.t | `Manual ]

  type encoder = {
    h : Dcmp.def; src : src; dst : dst;
    mutable i : bigstring; mutable o : bigstring; mutable o_len : int;
    mutable string_fed : bool;
  }
  type ret = [ `Await of encoder | `End of encoder | `Flush of encoder ]

  let empty = Dcmp.empty
  let io () = Bigarray.Array1.create Bigarray.char Bigarray.c_layout io_buffer_size

  (* zl.mli:170-186.  ?dynamic, ~q, ~w: the serial encoder's block choice and scratch; accepted, not used *)
  let encoder ?dynamic:(_ = true) ~q:(_ : De.Queue.t) ~w:(_ : De.Lz77.window) ~level src dst =
    let e = { h = Dcmp.def_new Dcmp.fmt_zlib level; src; dst; i = empty; o = empty; o_len = 0; string_fed = false } in
    (match dst with `Manual -> () | `Buffer _ | `Channel _ -> e.o <- io () ; e.o_len <- io_buffer_size ; Dcmp.def_dst e.h e.o 0 io_buffer_size)
    ; e

  let src e s j l = e.i <- s ; Dcmp.def_src e.h s j l ; e
  let dst e s j l = e.o <- s ; e.o_len <- l ; Dcmp.def_dst e.h s j l ; e
  let src_rem e = Dcmp.def_src_rem e.h
  let dst_rem e = Dcmp.def_dst_rem e.h

  let rec encode e : ret =                          (* was lib/zl.ml:509-592 *)
    match Dcmp.def_encode e.h with
    | 2 -> `End e
    | 1 -> (
      let len = e.o_len - Dcmp.def_dst_rem e.h in
      match e.dst with
      | `Manual -> `Flush e
      | `Buffer b -> for k = 0 to len - 1 do Buffer.add_char b (Bigarray.Array1.unsafe_get e.o k) done ; Dcmp.def_dst e.h e.o 0 e.o_len ; encode e
      | `Channel oc -> for k = 0 to len - 1 do output_char oc (Bigarray.Array1.unsafe_get e.o k) done ; Dcmp.def_dst e.h e.o 0 e.o_len ; encode e)
    | _ -> (
      match e.src with
      | `Manual -> `Await e
      | `String s ->
        if e.string_fed then Dcmp.def_src e.h empty 0 0
        else (e.string_fed <- true ; let b = Dcmp.bigstring_of_string s in e.i <- b ; Dcmp.def_src e.h b 0 (Bigarray.Array1.dim b))
        ; encode e
      | `Channel ic ->
        let buf = io () in
        let tmp = Bytes.create io_buffer_size in
        let n = input ic tmp 0 io_buffer_size in
        for k = 0 to n - 1 do Bigarray.Array1.unsafe_set buf k (Bytes.unsafe_get tmp k) done
        ; e.i <- buf ; Dcmp.def_src e.h buf 0 n ; encode e)

  module Ns = struct
    type error = De.Def.Ns.error
    let pp_error = De.Def.Ns.pp_error
    let compress_bound len = De.Def.Ns.compress_bound len + 6      (* zl.ml:598 *)
    let deflate ?(level = 4) src dst =              (* was lib/zl.ml:610-630 *)
      match Dcmp.def_ns_deflate Dcmp.fmt_zlib level src dst with
      | 0, len -> Ok len
      | 32, _ -> Error `Invalid_compression_level
      | _ -> Error `Unexpected_end_of_output
  end
end

module Higher = struct
  let compress ?(level = 6) ?dynamic ~w ~q ~refill ~flush i o =    (* was lib/zl.ml:634-648 *)
    let e = Def.encoder ?dynamic ~q ~w ~level `Manual `Manual in
    let rec go e =
      match Def.encode e with
      | `Await e -> let n = refill i in go (Def.src e i 0 n)
      | `Flush e -> flush o (Bigarray.Array1.dim o - Def.dst_rem e) ; go (Def.dst e o 0 (Bigarray.Array1.dim o))
      | `End _ -> () in
    go (Def.dst e o 0 (Bigarray.Array1.dim o))

  let uncompress ~allocate ~refill ~flush i o =     (* was lib/zl.ml:650-666 *)
    let rec go d =
      match Inf.decode d with
      | `Await d -> let n = refill i in go (Inf.src d i 0 n)
      | `Flush d -> flush o (Bigarray.Array1.dim o - Inf.dst_rem d) ; go (Inf.flush d)
      | `End d ->
        let len = Bigarray.Array1.dim o - Inf.dst_rem d in
        if len > 0 then flush o len ; Ok ()
      | `Malformed err -> Error (`Msg err) in
    go (Inf.decoder `Manual ~o ~allocate)
end
