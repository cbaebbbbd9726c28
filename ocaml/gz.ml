(* ocaml/gz.ml — drop-in body for lib/gz.ml behind the UNCHANGED lib/gz.mli: Gz.Inf / Gz.Def / Gz.Higher over
 * libdcmp_b200.so (dcmp_inf_* / dcmp_def_* with DCMP_FMT_GZIP).  The header quirks of the reference are kept by the C
 * side (csrc/stream_core.h): MTIME and the FEXTRA lengths big-endian (gz.ml:455, 476, 628), FHCRC = upper half of the
 * CRC-32 of the fixed bytes + FNAME + FCOMMENT (gz.ml:413-444). *)
type bigstring = De.bigstring
type window = De.window

let io_buffer_size = De.io_buffer_size

type os = FAT | Amiga | VMS | Unix | VM | Atari | HPFS | Macintosh | Z | CPM | TOPS20 | NTFS | QDOS | Acorn | Unknown

let os_table = [| FAT; Amiga; VMS; Unix; VM; Atari; HPFS; Macintosh; Z; CPM; TOPS20; NTFS; QDOS; Acorn |]
let os_of_int v = if v >= 0 && v < Array.length os_table then os_table.(v) else Unknown    (* gz.ml:231-246 *)
let int_of_os o = let r = ref 255 in Array.iteri (fun k x -> if x = o then r := k) os_table ; !r
let equal_os a b = a = b
let pp_os ppf o =
  Format.pp_print_string ppf
    (match o with FAT -> "FAT" | Amiga -> "Amiga" | VMS -> "VMS" | Unix -> "Unix" | VM -> "VM" | Atari -> "Atari"
     | HPFS -> "HPFS" | Macintosh -> "Macintosh" | Z -> "Z" | CPM -> "CPM" | TOPS20 -> "TOPS20" | NTFS -> "NTFS"
     | QDOS -> "QDOS" | Acorn -> "Acorn" | Unknown -> "Unknown")

module Inf = struct
  type src = Dcmp.src
  type decoder = Dcmp.decoder
  type signal = [ `Await of decoder | `Flush of decoder | `End of decoder | `Malformed of string ]

  let decoder (src : src) ~o : decoder = Dcmp.decoder Dcmp.fmt_gzip src ~o
  let decode (d : decoder) : signal =               (* was lib/gz.ml:494-531; `End after ONE member *)
    match Dcmp.decode d with
    | `Await -> `Await d | `Flush -> `Flush d | `End -> `End d | `Malformed err -> `Malformed err
  let reset d = Dcmp.reset d ; d                    (* then feed the src_rem bytes that follow the member (gz.ml:550) *)
  let src d s j l = Dcmp.src d s j l ; d
  let flush (d : decoder) = Dcmp.inf_flush d.Dcmp.h ; d
  let dst_rem (d : decoder) = Dcmp.inf_dst_rem d.Dcmp.h
  let src_rem (d : decoder) = Dcmp.inf_src_rem d.Dcmp.h
  let write (d : decoder) = Dcmp.inf_write d.Dcmp.h
  let filename (d : decoder) = Dcmp.gz_inf_field d.Dcmp.h 0
  let comment (d : decoder) = Dcmp.gz_inf_field d.Dcmp.h 1
  let os (d : decoder) = os_of_int (Dcmp.gz_inf_os d.Dcmp.h)

  let extra ~key (d : decoder) =                    (* gz.ml:617-633: 2-byte key, big-endian 16-bit length *)
    if String.length key <> 2 then invalid_arg "Subfield ID must be 2 characters."
    ; match Dcmp.gz_inf_field d.Dcmp.h 2 with
      | None -> None
      | Some payload ->
        let rec go idx =
          if idx + 4 > String.length payload then None
          else
            let len = (Char.code payload.[idx + 2] lsl 8) lor Char.code payload.[idx + 3] in
            if idx + 4 + len > String.length payload then None
            else if String.sub payload idx 2 = key then Some (String.sub payload (idx + 4) len)
            else go (idx + 4 + len) in
        go 0
end

module Def = struct
  type src = [ `Channel of in_channel | `String of string | `Manual ]
  type dst = [ `Channel of out_channel | `Buffer of Buffer.t | `Manual ]
  type encoder = Zl.Def.encoder
  type ret = [ `Await of encoder | `End of encoder | `Flush of encoder ]

  (* gz.mli:177-223 *)
  let encoder src dst ?(ascii = false) ?(hcrc = false) ?filename ?comment ~mtime os ~q:(_ : De.Queue.t)
      ~w:(_ : De.Lz77.window) ~level : encoder =
    let h = Dcmp.def_new Dcmp.fmt_gzip level in
    Dcmp.def_gz_header h mtime (int_of_os os) ((if ascii then 1 else 0) lor if hcrc then 2 else 0) filename comment
    ; let e = { Zl.Def.h; src; dst; i = Dcmp.empty; o = Dcmp.empty; o_len = 0; string_fed = false } in
      (match dst with
       | `Manual -> ()
       | `Buffer _ | `Channel _ -> e.Zl.Def.o <- Zl.Def.io () ; e.Zl.Def.o_len <- io_buffer_size ; Dcmp.def_dst h e.Zl.Def.o 0 io_buffer_size)
      ; e

  let src = Zl.Def.src
  let dst = Zl.Def.dst
  let src_rem = Zl.Def.src_rem
  let dst_rem = Zl.Def.dst_rem
  let encode : encoder -> ret = Zl.Def.encode       (* was lib/gz.ml:794-845 *)
end

module Higher = struct
  type 't configuration = { ascii : bool; hcrc : bool; os : os; mtime : 't -> int32 }
  let configuration ?(ascii = false) ?(hcrc = false) os mtime = { ascii; hcrc; os; mtime }

  let compress ?(level = 0) ?filename ?comment ~w ~q ~refill ~flush time cfg i o =    (* was lib/gz.ml:902-930 *)
    let e = Def.encoder `Manual `Manual ~ascii:cfg.ascii ~hcrc:cfg.hcrc ?filename ?comment ~mtime:(cfg.mtime time) cfg.os ~q ~w ~level in
    let rec go e =
      match Def.encode e with
      | `Await e -> let n = refill i in go (Def.src e i 0 n)
      | `Flush e -> flush o (Bigarray.Array1.dim o - Def.dst_rem e) ; go (Def.dst e o 0 (Bigarray.Array1.dim o))
      | `End _ -> () in
    go (Def.dst e o 0 (Bigarray.Array1.dim o))

  type metadata = { filename : string option; comment : string option; os : os; extra : key:string -> string option }

  let uncompress ~refill ~flush i o =               (* was lib/gz.ml:932-987 *)
    let rec go d =
      match Inf.decode d with
      | `Await d -> let n = refill i in go (Inf.src d i 0 n)
      | `Flush d -> flush o (Bigarray.Array1.dim o - Inf.dst_rem d) ; go (Inf.flush d)
      | `End d ->
        let len = Bigarray.Array1.dim o - Inf.dst_rem d in
        if len > 0 then flush o len
        ; Ok { filename = Inf.filename d; comment = Inf.comment d; os = Inf.os d; extra = (fun ~key -> Inf.extra ~key d) }
      | `Malformed err -> Error (`Msg err) in
    go (Inf.decoder `Manual ~o)
end
