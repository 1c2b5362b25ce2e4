"""Writes include/chase_fortran_interface.f90: the Fortran module `chase_diag` (same module and generic names as the
reference's interface/chase_fortran_interface.f90, so `use chase_diag` keeps working) as bind(C) declarations of the
entry points exported by libchase_b200.so.  Declarations only: no Fortran compiler is needed to build the library,
the application compiles this file with its own compiler.

    python scripts/gen_fortran_module.py
"""
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TYPES = {
    "s": ("REAL(c_float)", "REAL(c_float)", "c_float"),
    "d": ("REAL(c_double)", "REAL(c_double)", "c_double"),
    "c": ("COMPLEX(c_float_complex)", "REAL(c_float)", "c_float, c_float_complex"),
    "z": ("COMPLEX(c_double_complex)", "REAL(c_double)", "c_double, c_double_complex"),
}
out, generics = [], []


def sub(name, cname, args, decls, kinds):
    head = f"        SUBROUTINE {name}({', '.join(args)}) &"
    out.append(head)
    out.append(f"            bind(c, name='{cname}')")
    out.append(f"            USE, INTRINSIC :: iso_c_binding, ONLY: {kinds}")
    out.append("            IMPLICIT NONE")
    for d in decls:
        out.append("            " + d)
    out.append(f"        END SUBROUTINE {name}")
    out.append("")


def generic(gname, specifics):
    generics.append((gname, specifics))


for x, (vt, rt, kinds) in TYPES.items():
    k = "c_int, c_char, " + kinds
    seq_args = ["n", "nev", "nex", "h", "ldh", "v", "ritzv", "init"]
    seq_decl = ["INTEGER(c_int) :: n, nev, nex, ldh, init", f"{vt} :: h(ldh, *), v(n, *)", f"{rt} :: ritzv(*)"]
    int_args = ["n", "nev", "nex", "h", "ldh", "init"]
    int_decl = ["INTEGER(c_int) :: n, nev, nex, ldh, init", f"{vt} :: h(ldh, *)"]
    sub(f"{x}chase_init_with_buffers", f"{x}chase_init_", seq_args, seq_decl, k)
    sub(f"{x}chase_init_internal", f"{x}chase_init_internal_", int_args, int_decl, k)
    generic(f"{x}chase_init", [f"{x}chase_init_internal", f"{x}chase_init_with_buffers"])
    sub(f"{x}chase_finalize", f"{x}chase_finalize_", ["flag"], ["INTEGER(c_int) :: flag"], k)
    sub(f"{x}chase", f"{x}chase_", ["deg", "tol", "mode", "opt", "qr"],
        ["INTEGER(c_int) :: deg", f"{rt} :: tol", "CHARACTER(len=1, kind=c_char) :: mode, opt, qr"], k)
    sub(f"{x}chase_get_eigenpairs", f"{x}chase_get_eigenpairs_", ["v", "ld", "ritzv"],
        ["INTEGER(c_int) :: ld", f"{vt} :: v(ld, *)", f"{rt} :: ritzv(*)"], k)
    sub(f"{x}chase_readHam", f"{x}chase_readHam_", ["filename"], ["CHARACTER(kind=c_char) :: filename(*)"], k)
    if x in "cz":
        sub(f"{x}chase_init_pseudo_with_buffers", f"{x}chase_init_pseudo_f_", seq_args, seq_decl, k)
        sub(f"{x}chase_init_pseudo_internal", f"{x}chase_init_pseudo_internal_", int_args, int_decl, k)
        generic(f"{x}chase_init_pseudo", [f"{x}chase_init_pseudo_internal", f"{x}chase_init_pseudo_with_buffers"])
        sub(f"{x}chase_pseudo", f"{x}chase_pseudo_f_", ["deg", "tol", "mode", "opt", "qr"],
            ["INTEGER(c_int) :: deg", f"{rt} :: tol", "CHARACTER(len=1, kind=c_char) :: mode, opt, qr"], k)

    # distributed: one process per GPU; fcomm = chase_b200_comm_c2f(handle of chase_b200_comm_init)
    for ps in ("", "pseudo_") if x in "cz" else ("",):
        base = f"p{x}chase_init_{ps}"
        gbase = f"p{x}chase_init" + ("_pseudo" if ps else "")
        a = ["nn", "nev", "nex", "m", "n", "h", "ldh", "v", "ritzv", "dim0", "dim1", "grid_major", "fcomm", "init"]
        d = ["INTEGER(c_int) :: nn, nev, nex, m, n, ldh, dim0, dim1, fcomm, init", f"{vt} :: h(ldh, *), v(m, *)",
             f"{rt} :: ritzv(*)", "CHARACTER(len=1, kind=c_char) :: grid_major"]
        sub(f"{gbase}_with_buffers", f"{base}f_", a, d, k)
        a2 = [q for q in a if q not in ("v", "ritzv")]
        d2 = ["INTEGER(c_int) :: nn, nev, nex, m, n, ldh, dim0, dim1, fcomm, init", f"{vt} :: h(ldh, *)",
              "CHARACTER(len=1, kind=c_char) :: grid_major"]
        sub(f"{gbase}_internal", f"{base}internal_f_", a2, d2, k)
        generic(gbase, [f"{gbase}_internal", f"{gbase}_with_buffers"])
        a = ["nn", "nev", "nex", "mbsize", "nbsize", "h", "ldh", "v", "ritzv", "dim0", "dim1", "grid_major", "irsrc",
             "icsrc", "fcomm", "init"]
        d = ["INTEGER(c_int) :: nn, nev, nex, mbsize, nbsize, ldh, dim0, dim1, irsrc, icsrc, fcomm, init",
             f"{vt} :: h(ldh, *), v(ldh, *)", f"{rt} :: ritzv(*)", "CHARACTER(len=1, kind=c_char) :: grid_major"]
        sub(f"{gbase}_blockcyclic_with_buffers", f"{base}blockcyclic_f_", a, d, k)
        a2 = [q for q in a if q not in ("v", "ritzv")]
        d2 = ["INTEGER(c_int) :: nn, nev, nex, mbsize, nbsize, ldh, dim0, dim1, irsrc, icsrc, fcomm, init",
              f"{vt} :: h(ldh, *)", "CHARACTER(len=1, kind=c_char) :: grid_major"]
        sub(f"{gbase}_blockcyclic_internal", f"{base}blockcyclic_internal_f_", a2, d2, k)
        generic(f"{gbase}_blockcyclic", [f"{gbase}_blockcyclic_internal", f"{gbase}_blockcyclic_with_buffers"])
    sub(f"p{x}chase_finalize", f"p{x}chase_finalize_", ["flag"], ["INTEGER(c_int) :: flag"], k)
    sub(f"p{x}chase", f"p{x}chase_", ["deg", "tol", "mode", "opt", "qr"],
        ["INTEGER(c_int) :: deg", f"{rt} :: tol", "CHARACTER(len=1, kind=c_char) :: mode, opt, qr"], k)
    sub(f"p{x}chase_get_eigenpairs", f"p{x}chase_get_eigenpairs_", ["v", "ld", "ritzv"],
        ["INTEGER(c_int) :: ld", f"{vt} :: v(ld, *)", f"{rt} :: ritzv(*)"], k)
    sub(f"p{x}chase_wrtHam", f"p{x}chase_wrtHam_", ["filename"], ["CHARACTER(kind=c_char) :: filename(*)"], k)
    sub(f"p{x}chase_readHam", f"p{x}chase_readHam_", ["filename"], ["CHARACTER(kind=c_char) :: filename(*)"], k)

for nm, ft, kd in [("tol", "REAL(c_double)", "c_double"), ("decaying_rate", "REAL(c_float)", "c_float"),
                   ("upperb_scale_rate", "REAL(c_float)", "c_float")]:
    sub(f"chase_set_{nm}", f"chase_set_{nm}_", ["val"], [f"{ft} :: val"], kd)
for nm in ["deg", "max_deg", "deg_extra", "max_iter", "lanczos_iter", "num_lanczos", "approx", "opt", "cholqr",
           "cluster_aware_degrees"]:
    sub(f"chase_set_{nm}", f"chase_set_{nm}_", ["val"], ["INTEGER(c_int) :: val"], "c_int")
sub("chase_enable_sym_check", "chase_enable_sym_check_", ["flag"], ["INTEGER(c_int) :: flag"], "c_int")
for nm in ["cuda", "nccl", "scalapack", "mpi"]:
    sub(f"chase_has_{nm}", f"chase_has_{nm}_", ["flag"], ["INTEGER(c_int) :: flag"], "c_int")
sub("chase_get_version", "chase_get_version_", ["version", "length"],
    ["CHARACTER(kind=c_char) :: version(*)", "INTEGER(c_int) :: length"], "c_int, c_char")
sub("chase_print_config", "chase_print_config_", [], [], "c_int")

head = [
    "! chase_b200 -- Fortran module for libchase_b200.so (GENERATED by scripts/gen_fortran_module.py; do not edit).",
    "!",
    "! Module name, generic names and argument lists follow the reference's interface/chase_fortran_interface.f90",
    "! (MODULE chase_diag): an application written against the reference only has to be recompiled against this file and",
    "! linked with libchase_b200.so.  Every procedure is a bind(C) declaration of an entry point of",
    "! include/chase_c_interface.h (sequential: ?chase_init_, ?chase_, ...; distributed: the `_f_` twins that take the",
    "! communicator as an INTEGER, reference chase_c_interface.cpp:2296-2327, 2425-2900).",
    "!",
    "! Distributed runs are MPI-free here (one process per GPU, NCCL on the data path).  The INTEGER communicator is made by",
    "!     call chase_b200_comm_unique_id(id)            ! rank 0; ship the 128 bytes to every rank (MPI_Bcast, file, ...)",
    "!     ierr  = chase_b200_comm_init(rank, nranks, id, device, comm)",
    "!     fcomm = chase_b200_comm_c2f(comm)",
    "MODULE chase_diag",
    "    IMPLICIT NONE",
    "    PUBLIC",
    "",
]
for g, sp in generics:
    head.append(f"    INTERFACE {g}")
    for s_ in sp:
        head.append(f"        PROCEDURE :: {s_}")
    head.append(f"    END INTERFACE {g}")
    head.append("")
body = ["    INTERFACE", ""] + out + [
    "        ! communicator bootstrap (include/chase_b200_comm.h)",
    "        INTEGER(c_int) FUNCTION chase_b200_comm_unique_id(id) bind(c, name='chase_b200_comm_unique_id')",
    "            USE, INTRINSIC :: iso_c_binding, ONLY: c_int, c_signed_char",
    "            INTEGER(c_signed_char) :: id(128)",
    "        END FUNCTION chase_b200_comm_unique_id",
    "        INTEGER(c_int) FUNCTION chase_b200_comm_init(rank, nranks, id, device, comm) &",
    "            bind(c, name='chase_b200_comm_init')",
    "            USE, INTRINSIC :: iso_c_binding, ONLY: c_int, c_signed_char, c_ptr",
    "            INTEGER(c_int), VALUE :: rank, nranks, device",
    "            INTEGER(c_signed_char) :: id(128)",
    "            TYPE(c_ptr) :: comm",
    "        END FUNCTION chase_b200_comm_init",
    "        INTEGER(c_int) FUNCTION chase_b200_comm_c2f(comm) bind(c, name='chase_b200_comm_c2f')",
    "            USE, INTRINSIC :: iso_c_binding, ONLY: c_int, c_ptr",
    "            TYPE(c_ptr), VALUE :: comm",
    "        END FUNCTION chase_b200_comm_c2f",
    "        INTEGER(c_int) FUNCTION chase_b200_comm_free(comm) bind(c, name='chase_b200_comm_free')",
    "            USE, INTRINSIC :: iso_c_binding, ONLY: c_int, c_ptr",
    "            TYPE(c_ptr), VALUE :: comm",
    "        END FUNCTION chase_b200_comm_free",
    "",
    "    END INTERFACE",
    "END MODULE chase_diag",
]
path = os.path.join(ROOT, "include", "chase_fortran_interface.f90")
open(path, "w").write("\n".join(head + body) + "\n")
print(path, len(head) + len(body), "lines;", sum(1 for l in out if "bind(c" in l), "bind(C) procedures")
