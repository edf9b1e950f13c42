//! Raw bindings of include/kanpyo_b200.h (ABI version 1).  SOURCE ONLY: the build image has no Rust
//! toolchain, so this file is checked by review against the header, not by a compiler.  The same
//! symbols are exercised from C++ (tests/cpp/test_tokenizer.cpp) and Python (kanpyo_b200/_lib.py).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int};

pub const KP_OK: c_int = 0;

#[repr(C)]
pub struct kp_dict_arrays {
    pub da: *const i32,
    pub da_len: u64,
    pub dup_ids: *const i64,
    pub dup_counts: *const u64,
    pub n_dup: u64,
    pub morphs: *const i16,
    pub n_morphs: u64,
    pub conn_row: u64,
    pub conn_col: u64,
    pub conn: *const i16,
    pub char_category: *const u8,
    pub n_char_category: u64,
    pub invoke_list: *const u8,
    pub n_invoke: u64,
    pub group_list: *const u8,
    pub n_group: u64,
    pub unk_cat: *const u8,
    pub unk_first_id: *const i64,
    pub unk_count: *const u64,
    pub n_unk_map: u64,
    pub unk_morphs: *const i16,
    pub n_unk_morphs: u64,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct kp_token {
    pub id: i32,
    pub position: u32,
    pub start: u32,
    pub char_len: u16,
    pub cls: u8,
    pub reserved: u8,
}

#[repr(C)]
pub struct kp_result {
    pub n_sent: u64,
    pub n_tokens: u64,
    pub tok_off: *const u64,
    pub tokens: *const kp_token,
    pub eos_cost: *const i32,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct kp_lattice_node {
    pub id: i32,
    pub cls: u8,
    pub reserved: [u8; 3],
    pub byte_pos: u32,
    pub char_pos: u32,
    pub end_char: u32,
    pub left_id: i16,
    pub right_id: i16,
    pub cost: i16,
    pub reserved2: i16,
    pub dp: i32,
    pub pre: i32,
}

#[repr(C)]
pub struct kp_lattice {
    pub n_nodes: u64,
    pub nodes: *const kp_lattice_node,
}

pub enum kp_dict {}
pub enum kp_tokenizer {}

#[link(name = "kanpyo_b200")]
extern "C" {
    pub fn kp_abi_version() -> c_int;
    pub fn kp_strerror(status: c_int) -> *const c_char;
    pub fn kp_last_error() -> *const c_char;
    pub fn kp_dict_create(arrays: *const kp_dict_arrays, device: c_int, out: *mut *mut kp_dict) -> c_int;
    pub fn kp_dict_destroy(d: *mut kp_dict);
    pub fn kp_tokenizer_create(d: *const kp_dict, out: *mut *mut kp_tokenizer) -> c_int;
    pub fn kp_tokenizer_destroy(t: *mut kp_tokenizer);
    pub fn kp_tokenize(t: *mut kp_tokenizer, utf8: *const u8, len: u64, out: *mut kp_result) -> c_int;
    pub fn kp_tokenize_batch(
        t: *mut kp_tokenizer,
        utf8: *const u8,
        offsets: *const u64,
        n_sent: u64,
        out: *mut kp_result,
    ) -> c_int;
    pub fn kp_lattice_dump(t: *mut kp_tokenizer, utf8: *const u8, len: u64, out: *mut kp_lattice) -> c_int;
}
