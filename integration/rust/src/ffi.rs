//! Raw bindings of include/kanpyo_b200.h (ABI version 2).  SOURCE ONLY: the build image has no Rust
//! toolchain, so this file is checked by review against the header, not by a compiler.  The same
//! symbols are exercised from C++ (tests/cpp/test_tokenizer.cpp) and Python (kanpyo_b200/_lib.py).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int};

pub const KP_OK: c_int = 0;
pub const KP_PATH_AUTO: c_int = 0;
pub const KP_PATH_PIPELINE: c_int = 1;
pub const KP_PATH_FUSED: c_int = 2;

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

/// `Token` without the heap string (src/token.rs:11-18).
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

/// Compact transfer record: half the device-to-host bytes of `kp_token`.  `position` / `start` are
/// rebuilt on the host by walking a sentence's records backwards from its EOS record, which carries
/// the sentence's char count in its two length fields (see the header).
#[repr(C)]
#[derive(Clone, Copy)]
pub struct kp_token8 {
    pub id_cls: u32,
    pub byte_len: u16,
    pub char_len: u16,
}

#[repr(C)]
pub struct kp_result8 {
    pub n_sent: u64,
    pub n_tokens: u64,
    pub tok_off: *const u32,
    pub tokens: *const kp_token8,
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
pub enum kp_queue {}
pub enum kp_shards {}

#[link(name = "kanpyo_b200")]
extern "C" {
    pub fn kp_abi_version() -> c_int;
    pub fn kp_strerror(status: c_int) -> *const c_char;
    pub fn kp_last_error() -> *const c_char;
    pub fn kp_device_count(n: *mut c_int) -> c_int;
    pub fn kp_dict_create(arrays: *const kp_dict_arrays, device: c_int, out: *mut *mut kp_dict) -> c_int;
    pub fn kp_dict_destroy(d: *mut kp_dict);
    pub fn kp_tokenizer_create(d: *const kp_dict, out: *mut *mut kp_tokenizer) -> c_int;
    pub fn kp_tokenizer_destroy(t: *mut kp_tokenizer);
    pub fn kp_tokenizer_set_path(t: *mut kp_tokenizer, path: c_int) -> c_int;
    pub fn kp_tokenize(t: *mut kp_tokenizer, utf8: *const u8, len: u64, out: *mut kp_result) -> c_int;
    pub fn kp_tokenize_batch(
        t: *mut kp_tokenizer,
        utf8: *const u8,
        offsets: *const u64,
        n_sent: u64,
        out: *mut kp_result,
    ) -> c_int;
    pub fn kp_tokenize_batch8(
        t: *mut kp_tokenizer,
        utf8: *const u8,
        offsets: *const u64,
        n_sent: u64,
        out: *mut kp_result8,
    ) -> c_int;
    pub fn kp_expand_tokens8(r: *const kp_result8, offsets: *const u64, out: *mut kp_token) -> c_int;
    pub fn kp_lattice_dump(t: *mut kp_tokenizer, utf8: *const u8, len: u64, out: *mut kp_lattice) -> c_int;
    // successive batches with their copies hidden behind each other's kernels
    pub fn kp_queue_create(d: *const kp_dict, depth: u32, out: *mut *mut kp_queue) -> c_int;
    pub fn kp_queue_submit(q: *mut kp_queue, utf8: *const u8, offsets: *const u64, n_sent: u64, ticket: *mut u64) -> c_int;
    pub fn kp_queue_wait(q: *mut kp_queue, ticket: u64, out: *mut kp_result8) -> c_int;
    pub fn kp_queue_destroy(q: *mut kp_queue);
    // every GPU of the box from this one process: NCCL dictionary broadcast, byte-balanced sentence shards
    pub fn kp_shards_create(arrays: *const kp_dict_arrays, devices: *const c_int, n_devices: c_int, out: *mut *mut kp_shards) -> c_int;
    pub fn kp_shards_tokenize(g: *mut kp_shards, utf8: *const u8, offsets: *const u64, n_sent: u64, out: *mut kp_result8) -> c_int;
    pub fn kp_shards_destroy(g: *mut kp_shards);
}
