"""Shared test helpers: conversion of the product's host containers to the oracle's raw containers and
oracle dispatch."""
import numpy as np

import bsm_b200 as B
from oracle import oracle_np as O


def to_oracle(A):
    if isinstance(A, B.SymmetricBlockMatrix):
        return O.OSBM(A.diagonals, A.diagonalindices, A.offdiagonals, A.rowindices, A.colindices, A.size)
    if isinstance(A, B.VariableBlockCompressedRowStorage):
        return O.OVBCRS(A.blocks, A.rowptr, A.colindices, A.rowindices, A.size)
    return O.OBSM(A.blocks, A.rowindices, A.colindices, A.size)


def oracle_mul(A, x, op="N", alpha=1, beta=0, beta_is_false=True, y=None, threads=4, f64=False):
    """C oracle product. f64=True evaluates a Float32 matrix in Float64 (for the 1e-5 bound)."""
    OA = to_oracle(A)
    if f64:
        up = lambda bs: [np.asfortranarray(b, dtype=np.float64) for b in bs]
        if isinstance(OA, O.OSBM):
            OA = O.OSBM(up(OA.diagonals), OA.diagonalindices, up(OA.offdiagonals), OA.rowindices, OA.colindices, OA.size)
        elif isinstance(OA, O.OVBCRS):
            OA = O.OVBCRS(up(OA.blocks), OA.rowptr, OA.colindices, OA.rowindices, OA.size)
        else:
            OA = O.OBSM(up(OA.blocks), OA.rowindices, OA.colindices, OA.size)
        x = np.asarray(x, np.float64)
        y = None if y is None else np.asarray(y, np.float64)
    if isinstance(OA, O.OSBM):
        return O.c_mul_sbm(OA, x, op, alpha, beta, beta_is_false, y, threads)
    if isinstance(OA, O.OVBCRS):
        return O.c_mul_vbcrs(OA, x, op, alpha, beta, beta_is_false, y, threads)
    return O.c_mul_bsm(OA, x, op, alpha, beta, beta_is_false, y, threads)


def rel2(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    return float(np.linalg.norm(a.astype(np.complex128) - b) / max(np.linalg.norm(b), 1e-300))


def randx(rng, n, dtype):
    dtype = np.dtype(dtype)
    x = rng.standard_normal(n)
    if dtype.kind == "c":
        x = x + 1j * rng.standard_normal(n)
    return x.astype(dtype)


TOL = {np.dtype(np.float32): 1e-5, np.dtype(np.float64): 1e-12, np.dtype(np.complex128): 1e-12}
