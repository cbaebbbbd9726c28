(* ocaml/lzo.ml — drop-in body for the decode half of lib/lzo.ml behind the UNCHANGED lib/lzo.mli; Lzo.compress and
 * make_wrkmem (lzo.ml:416-656) are not on the hot path and stay the reference's (Lzo_ref = the renamed original). *)
include Lzo_ref

type bigstring = (char, Bigarray.int8_unsigned_elt, Bigarray.c_layout) Bigarray.Array1.t
type error = [ `Malformed of string | `Invalid_argument of string | `Invalid_dictionary ]

let error_of_status : int -> error = function       (* include/dcmp.h: 16..18 = Lzo.error (lzo.ml:4-12) *)
  | 16 -> `Malformed "Malformed input"
  | 18 -> `Invalid_dictionary
  | _ -> `Invalid_argument "Input is malformed or output is not large enough"

let uncompress input output =                       (* was lib/lzo.ml:395-403 *)
  match Dcmp.lzo_uncompress 0 input output with
  | 0, len -> Ok (Bigarray.Array1.sub output 0 len)
  | st, _ -> Error (error_of_status st :> [> error ])

let uncompress_with_buffer ?(chunk = 0x1000) input = (* was lib/lzo.ml:405-414: the output grows until it fits *)
  let rec go cap =
    let output = Bigarray.Array1.create Bigarray.char Bigarray.c_layout cap in
    match Dcmp.lzo_uncompress 1 input output with
    | 0, len -> Ok (String.init len (Bigarray.Array1.get output))
    | 2, _ -> go (cap * 4)                           (* DCMP_UNEXPECTED_END_OF_OUTPUT: dst was too small *)
    | st, _ -> Error (error_of_status st :> [> error ]) in
  go (max chunk (4 * Bigarray.Array1.dim input))
