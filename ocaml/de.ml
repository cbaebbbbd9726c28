(* ocaml/de.ml — drop-in body for lib/de.ml behind the UNCHANGED lib/de.mli.
 *
 * Build recipe: rename the reference's lib/de.ml to lib/de_ref.ml (not exposed), add this file as lib/de.ml, dcmp.ml and
 * dcmp_stubs.c next to it (INTEGRATION.md §2).  Everything that is host bookkeeping in the reference — Queue, the
 * frequency counters, Lookup, T, the command-level Def.encoder and Lz77.state (de.mli:188-524) — comes from De_ref
 * untouched; what is the hot path — the inflate state machine (Inf, Inf.Ns), the one-shot deflater (Def.Ns) and the
 * Higher loops — is re-bound to libdcmp_b200.so here. *)
include De_ref

module Inf = struct
  type src = Dcmp.src
  type decode = [ `Await | `Flush | `End | `Malformed of string ]
  type decoder = Dcmp.decoder

  (* de.mli:96-116.  ~w: the window is a ring on the device; the caller's is accepted and left alone. *)
  let decoder (src : src) ~o ~w:(_ : window) : decoder = Dcmp.decoder Dcmp.fmt_raw src ~o
  let decode : decoder -> decode = Dcmp.decode                       (* was lib/de.ml:1468-1474 *)
  let reset = Dcmp.reset                                             (* lib/de.ml:1513-1532 *)
  let src = Dcmp.src                                                 (* lib/de.ml:733-742; Invalid_argument on bad bounds *)
  let dst_rem (d : decoder) = Dcmp.inf_dst_rem d.Dcmp.h
  let src_rem (d : decoder) = Dcmp.inf_src_rem d.Dcmp.h
  let flush (d : decoder) = Dcmp.inf_flush d.Dcmp.h
  let checksum (d : decoder) = Optint.of_unsigned_int32 (Dcmp.inf_checksum d.Dcmp.h)

  module Ns = struct
    type error =
      [ `Unexpected_end_of_input | `Unexpected_end_of_output | `Invalid_kind_of_block | `Invalid_dictionary
      | `Invalid_complement_of_length | `Invalid_distance | `Invalid_distance_code ]

    let pp_error = De_ref.Inf.Ns.pp_error

    let error_of_status : int -> error = function   (* include/dcmp.h: 1..7 in the order of de.mli:151-157 *)
      | 1 -> `Unexpected_end_of_input | 2 -> `Unexpected_end_of_output | 3 -> `Invalid_kind_of_block
      | 4 -> `Invalid_dictionary | 5 -> `Invalid_complement_of_length | 6 -> `Invalid_distance
      | 7 -> `Invalid_distance_code
      | _ -> invalid_arg "index out of bounds"   (* status 8: the pure-OCaml body raises there too (de.ml:640-647) *)

    let inflate src dst =                           (* was lib/de.ml:1807-1822 *)
      match Dcmp.inf_ns_inflate Dcmp.fmt_raw src dst with
      | 0, consumed, produced -> Ok (consumed, produced)
      | st, _, _ -> Error (error_of_status st :> [> error ])
  end
end

module Def = struct
  include De_ref.Def                                (* encoder / encode / dst at command level: unchanged *)

  module Ns = struct
    type error = [ `Invalid_compression_level | `Unexpected_end_of_output ]

    let pp_error = De_ref.Def.Ns.pp_error
    let compress_bound = De_ref.Def.Ns.compress_bound   (* de.ml:3993-3996; dcmp_deflate_bound is the same formula *)

    let deflate ?(level = 4) src dst =              (* was lib/de.ml:3998-4010 *)
      match Dcmp.def_ns_deflate Dcmp.fmt_raw level src dst with
      | 0, len -> Ok len
      | 32, _ -> Error `Invalid_compression_level
      | _ -> Error `Unexpected_end_of_output
  end
end

module Higher = struct
  (* de.mli:540-575: ~w and ~q are the serial matcher's scratch; the GPU path has its own on the device *)
  let compress ?(level = 4) ~w:(_ : Lz77.window) ~q:(_ : Queue.t) ~refill ~flush i o =
    let e = Dcmp.def_new Dcmp.fmt_raw level in
    Dcmp.def_dst e o 0 (Bigarray.Array1.dim o)
    ; let rec go () =
        match Dcmp.def_encode e with
        | 0 -> let n = refill i in Dcmp.def_src e i 0 n ; go ()
        | 1 -> flush o (Bigarray.Array1.dim o - Dcmp.def_dst_rem e) ; Dcmp.def_dst e o 0 (Bigarray.Array1.dim o) ; go ()
        | _ -> () in
      go ()

  let uncompress ~w ~refill ~flush i o =            (* was lib/de.ml:4562-4580 *)
    let d = Inf.decoder `Manual ~o ~w in
    let rec go () =
      match Inf.decode d with
      | `Await -> let n = refill i in Inf.src d i 0 n ; go ()
      | `Flush -> flush o (Bigarray.Array1.dim o - Inf.dst_rem d) ; Inf.flush d ; go ()
      | `End ->
        let len = Bigarray.Array1.dim o - Inf.dst_rem d in
        if len > 0 then flush o len ; Ok ()
      | `Malformed err -> Error (`Msg err) in
    go ()

  let of_string ~o ~w str ~flush =
    let d = Inf.decoder (`String str) ~o ~w in
    let rec go () =
      match Inf.decode d with
      | `Await -> assert false
      | `Flush -> flush o (Bigarray.Array1.dim o - Inf.dst_rem d) ; Inf.flush d ; go ()
      | `End ->
        let len = Bigarray.Array1.dim o - Inf.dst_rem d in
        if len > 0 then flush o len ; Ok ()
      | `Malformed err -> Error (`Msg err) in
    go ()

  let to_string ?(buffer = 4096) ~w ~q ~refill i =
    let buf = Buffer.create buffer in
    let o = Bigarray.Array1.create Bigarray.char Bigarray.c_layout io_buffer_size in
    let flush o len = for k = 0 to len - 1 do Buffer.add_char buf (Bigarray.Array1.unsafe_get o k) done in
    compress ~w ~q ~refill ~flush i o ; Buffer.contents buf
end
