//! Drop-in for the reference's `src/tokenizer.rs`: same `Tokenizer { pub dict }`, `Tokenizer::new`,
//! `Tokenizer::tokenize(&self, &str) -> Vec<Token>`, with the lattice build and Viterbi search running
//! in the CUDA library behind `ffi`.  Adds `tokenize_batch`.  SOURCE ONLY (no Rust toolchain in the
//! build image); see INTEGRATION.md.
use std::sync::Mutex;

use kanpyo_dict::dict::Dict;

use crate::ffi;
use crate::token::{Token, TokenClass};

/// Device-side state: the dictionary staged in HBM once and one tokenizer handle (one CUDA stream).
/// `kp_tokenizer` is not re-entrant, `tokenize(&self)` must be: the handle sits behind a mutex.
struct Device {
    dict: *mut ffi::kp_dict,
    tokenizer: Mutex<*mut ffi::kp_tokenizer>,
}
unsafe impl Send for Device {}
unsafe impl Sync for Device {}

impl Drop for Device {
    fn drop(&mut self) {
        unsafe {
            ffi::kp_tokenizer_destroy(*self.tokenizer.lock().unwrap());
            ffi::kp_dict_destroy(self.dict);
        }
    }
}

pub struct Tokenizer {
    pub dict: Dict,
    device: Device,
}

fn expect_ok(status: i32, what: &str) {
    if status != ffi::KP_OK {
        let detail = unsafe { std::ffi::CStr::from_ptr(ffi::kp_last_error()) };
        // the reference's tokenize() is infallible by signature and panics on bad indices
        panic!("kanpyo_b200: {} failed with status {}: {}", what, status, detail.to_string_lossy());
    }
}

impl Tokenizer {
    pub fn new(dict: Dict) -> Self {
        // flatten exactly the members the hot path reads (kanpyo-dict/src/dict.rs:21-30); the five
        // read-only accessors used here are listed in ../kanpyo_dict_accessors.rs
        let da: Vec<i32> = dict.index_table.da_nodes().iter().flat_map(|n| [n.base, n.check]).collect();
        let dup_ids: Vec<i64> = dict.index_table.dup_map().keys().map(|&k| k as i64).collect();
        let dup_counts: Vec<u64> = dict.index_table.dup_map().values().map(|&v| v as u64).collect();
        let morphs: Vec<i16> = dict.morphs.as_slice().iter().flat_map(|m| [m.left_id, m.right_id, m.cost]).collect();
        let unk_morphs: Vec<i16> =
            dict.unk_dict.morphs.as_slice().iter().flat_map(|m| [m.left_id, m.right_id, m.cost]).collect();
        let invoke: Vec<u8> = dict.char_category_def.invoke_list.iter().map(|&b| b as u8).collect();
        let group: Vec<u8> = dict.char_category_def.group_list.iter().map(|&b| b as u8).collect();
        let unk_cat: Vec<u8> = dict.unk_dict.char_category_to_morph_id.keys().copied().collect();
        let unk_first: Vec<i64> =
            dict.unk_dict.char_category_to_morph_id.values().map(|&(id, _)| id as i64).collect();
        let unk_count: Vec<u64> =
            dict.unk_dict.char_category_to_morph_id.values().map(|&(_, n)| n as u64).collect();
        let arrays = ffi::kp_dict_arrays {
            da: da.as_ptr(),
            da_len: (da.len() / 2) as u64,
            dup_ids: dup_ids.as_ptr(),
            dup_counts: dup_counts.as_ptr(),
            n_dup: dup_ids.len() as u64,
            morphs: morphs.as_ptr(),
            n_morphs: (morphs.len() / 3) as u64,
            conn_row: dict.connection_table.shape().0 as u64,
            conn_col: dict.connection_table.shape().1 as u64,
            conn: dict.connection_table.as_slice().as_ptr(),
            char_category: dict.char_category_def.char_category.as_ptr(),
            n_char_category: dict.char_category_def.char_category.len() as u64,
            invoke_list: invoke.as_ptr(),
            n_invoke: invoke.len() as u64,
            group_list: group.as_ptr(),
            n_group: group.len() as u64,
            unk_cat: unk_cat.as_ptr(),
            unk_first_id: unk_first.as_ptr(),
            unk_count: unk_count.as_ptr(),
            n_unk_map: unk_cat.len() as u64,
            unk_morphs: unk_morphs.as_ptr(),
            n_unk_morphs: (unk_morphs.len() / 3) as u64,
        };
        let mut d = std::ptr::null_mut();
        let mut t = std::ptr::null_mut();
        unsafe {
            expect_ok(ffi::kp_dict_create(&arrays, 0, &mut d), "kp_dict_create");
            expect_ok(ffi::kp_tokenizer_create(d, &mut t), "kp_tokenizer_create");
        }
        Self { dict, device: Device { dict: d, tokenizer: Mutex::new(t) } }
    }

    pub fn tokenize(&self, input: &str) -> Vec<Token> {
        self.tokenize_batch(&[input]).pop().unwrap()
    }

    /// `tokenize` over many independent sentences in one device pass.
    pub fn tokenize_batch(&self, inputs: &[&str]) -> Vec<Vec<Token>> {
        let mut text = Vec::with_capacity(inputs.iter().map(|s| s.len()).sum());
        let mut offsets = Vec::with_capacity(inputs.len() + 1);
        offsets.push(0u64);
        for s in inputs {
            text.extend_from_slice(s.as_bytes());
            offsets.push(text.len() as u64);
        }
        let handle = self.device.tokenizer.lock().unwrap();
        let mut r = std::mem::MaybeUninit::<ffi::kp_result>::uninit();
        let r = unsafe {
            expect_ok(
                ffi::kp_tokenize_batch(*handle, text.as_ptr(), offsets.as_ptr(), inputs.len() as u64, r.as_mut_ptr()),
                "kp_tokenize_batch",
            );
            r.assume_init()
        };
        let tok_off = unsafe { std::slice::from_raw_parts(r.tok_off, inputs.len() + 1) };
        let tokens = unsafe { std::slice::from_raw_parts(r.tokens, r.n_tokens as usize) };
        inputs
            .iter()
            .enumerate()
            .map(|(s, input)| {
                let toks = &tokens[tok_off[s] as usize..tok_off[s + 1] as usize];
                toks.iter()
                    .enumerate()
                    .map(|(k, t)| {
                        let class = match t.cls {
                            0 => TokenClass::Dummy,
                            1 => TokenClass::Known,
                            _ => TokenClass::Unknown,
                        };
                        // consecutive path nodes are adjacent in the input; EOS closes the path
                        let surface = if class == TokenClass::Dummy {
                            "EOS".to_string()
                        } else {
                            input[t.position as usize..toks[k + 1].position as usize].to_string()
                        };
                        Token {
                            id: t.id as isize,
                            class,
                            position: t.position as usize,
                            start: t.start as usize,
                            end: t.start as usize + t.char_len as usize,
                            surface,
                        }
                    })
                    .collect()
            })
            .collect()
    }
}
