"""kanpyo_b200 — B200-native (sm_100a) implementation of Kanpyo's tokenizer hot path.

Host-side mirror of the reference's public API for the path:

    reference (Rust)                          here
    ---------------------------------------   -----------------------------------------
    kanpyo_dict::dict::Dict                   kanpyo_b200.Dict
    kanpyo::tokenizer::Tokenizer::new(dict)   kanpyo_b200.Tokenizer(dict)
    Tokenizer::tokenize(&str) -> Vec<Token>   Tokenizer.tokenize(str) -> list[Token]
    kanpyo::token::{Token, TokenClass}        kanpyo_b200.Token, kanpyo_b200.TokenClass
    Lattice::build + Lattice::viterbi         Tokenizer.lattice(str)

All compute happens in hand-written CUDA kernels behind the C ABI of include/kanpyo_b200.h;
there is no CPU fallback.
"""
from .dict import Dict  # noqa: F401
from .tokenizer import Token, TokenClass, Tokenizer, BatchResult  # noqa: F401
from ._lib import KanpyoB200Error  # noqa: F401

__all__ = ["Dict", "Tokenizer", "Token", "TokenClass", "BatchResult", "KanpyoB200Error"]
