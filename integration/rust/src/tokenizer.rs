//! Drop-in for the reference's `src/tokenizer.rs`: same `Tokenizer { pub dict }`, `Tokenizer::new`,
//! `Tokenizer::tokenize(&self, &str) -> Vec<Token>`, with the lattice build and Viterbi search running
//! in the CUDA library behind `ffi`.  Adds `tokenize_batch` (one GPU) and `tokenize_batch_all_gpus`.
//! SOURCE ONLY (no Rust toolchain in the build image); see INTEGRATION.md.
use std::collections::HashMap;
use std::os::raw::c_int;
use std::sync::{Arc, Mutex, OnceLock, Weak};

use kanpyo_dict::dict::Dict;

use crate::ffi;
use crate::token::{Token, TokenClass};

/// Device-side state of one dictionary: the arrays staged in HBM once, a pool of tokenizer handles (a
/// `kp_tokenizer` is one CUDA stream + scratch and is not re-entrant; `tokenize(&self)` must be, so
/// every concurrent caller takes its own handle from the pool), and, once asked for, the multi-GPU group.
pub(crate) struct Device {
    dict: *mut ffi::kp_dict,
    pool: Mutex<Vec<*mut ffi::kp_tokenizer>>,
    shards: Mutex<Option<*mut ffi::kp_shards>>,
}
unsafe impl Send for Device {}
unsafe impl Sync for Device {}

impl Drop for Device {
    fn drop(&mut self) {
        unsafe {
            for t in self.pool.lock().unwrap().drain(..) {
                ffi::kp_tokenizer_destroy(t);
            }
            if let Some(g) = self.shards.lock().unwrap().take() {
                ffi::kp_shards_destroy(g);
            }
            ffi::kp_dict_destroy(self.dict);
        }
    }
}

fn expect_ok(status: i32, what: &str) {
    if status != ffi::KP_OK {
        let detail = unsafe { std::ffi::CStr::from_ptr(ffi::kp_last_error()) };
        // the reference's tokenize() is infallible by signature and panics on bad indices
        panic!("kanpyo_b200: {} failed with status {}: {}", what, status, detail.to_string_lossy());
    }
}

/// The flat views of the members the hot path reads (kanpyo-dict/src/dict.rs:21-30); the five
/// read-only accessors used here are listed in ../kanpyo_dict_accessors.rs.
struct Flat {
    da: Vec<i32>,
    dup_ids: Vec<i64>,
    dup_counts: Vec<u64>,
    morphs: Vec<i16>,
    unk_morphs: Vec<i16>,
    invoke: Vec<u8>,
    group: Vec<u8>,
    unk_cat: Vec<u8>,
    unk_first: Vec<i64>,
    unk_count: Vec<u64>,
}

impl Flat {
    fn of(dict: &Dict) -> Self {
        Flat {
            da: dict.index_table.da_nodes().iter().flat_map(|n| [n.base, n.check]).collect(),
            dup_ids: dict.index_table.dup_map().keys().map(|&k| k as i64).collect(),
            dup_counts: dict.index_table.dup_map().values().map(|&v| v as u64).collect(),
            morphs: dict.morphs.as_slice().iter().flat_map(|m| [m.left_id, m.right_id, m.cost]).collect(),
            unk_morphs: dict.unk_dict.morphs.as_slice().iter().flat_map(|m| [m.left_id, m.right_id, m.cost]).collect(),
            invoke: dict.char_category_def.invoke_list.iter().map(|&b| b as u8).collect(),
            group: dict.char_category_def.group_list.iter().map(|&b| b as u8).collect(),
            unk_cat: dict.unk_dict.char_category_to_morph_id.keys().copied().collect(),
            unk_first: dict.unk_dict.char_category_to_morph_id.values().map(|&(id, _)| id as i64).collect(),
            unk_count: dict.unk_dict.char_category_to_morph_id.values().map(|&(_, n)| n as u64).collect(),
        }
    }

    fn arrays(&self, dict: &Dict) -> ffi::kp_dict_arrays {
        ffi::kp_dict_arrays {
            da: self.da.as_ptr(),
            da_len: (self.da.len() / 2) as u64,
            dup_ids: self.dup_ids.as_ptr(),
            dup_counts: self.dup_counts.as_ptr(),
            n_dup: self.dup_ids.len() as u64,
            morphs: self.morphs.as_ptr(),
            n_morphs: (self.morphs.len() / 3) as u64,
            conn_row: dict.connection_table.shape().0 as u64,
            conn_col: dict.connection_table.shape().1 as u64,
            conn: dict.connection_table.as_slice().as_ptr(),
            char_category: dict.char_category_def.char_category.as_ptr(),
            n_char_category: dict.char_category_def.char_category.len() as u64,
            invoke_list: self.invoke.as_ptr(),
            n_invoke: self.invoke.len() as u64,
            group_list: self.group.as_ptr(),
            n_group: self.group.len() as u64,
            unk_cat: self.unk_cat.as_ptr(),
            unk_first_id: self.unk_first.as_ptr(),
            unk_count: self.unk_count.as_ptr(),
            n_unk_map: self.unk_cat.len() as u64,
            unk_morphs: self.unk_morphs.as_ptr(),
            n_unk_morphs: (self.unk_morphs.len() / 3) as u64,
        }
    }
}

/// Device state per `Dict`, found again from a bare `&Dict` (that is all `Lattice::build(&Dict, &str)`
/// gets, src/lattice.rs:101).  The key is the address of the morph table's heap buffer: it does not
/// change when the `Dict` (or the `Tokenizer` owning it) is moved.
fn registry() -> &'static Mutex<HashMap<usize, Weak<Device>>> {
    static R: OnceLock<Mutex<HashMap<usize, Weak<Device>>>> = OnceLock::new();
    R.get_or_init(|| Mutex::new(HashMap::new()))
}

pub(crate) fn device_of(dict: &Dict) -> Arc<Device> {
    let key = dict.morphs.as_slice().as_ptr() as usize;
    let mut reg = registry().lock().unwrap();
    if let Some(dev) = reg.get(&key).and_then(Weak::upgrade) {
        return dev;
    }
    let flat = Flat::of(dict);
    let arrays = flat.arrays(dict);
    let mut d = std::ptr::null_mut();
    unsafe { expect_ok(ffi::kp_dict_create(&arrays, 0, &mut d), "kp_dict_create") };
    let dev = Arc::new(Device { dict: d, pool: Mutex::new(vec![]), shards: Mutex::new(None) });
    reg.insert(key, Arc::downgrade(&dev));
    dev
}

impl Device {
    /// Runs `f` with a tokenizer handle of its own (created on first use, returned to the pool after).
    pub(crate) fn with_handle<R>(&self, f: impl FnOnce(*mut ffi::kp_tokenizer) -> R) -> R {
        let taken = self.pool.lock().unwrap().pop();
        let handle = taken.unwrap_or_else(|| {
            let mut t = std::ptr::null_mut();
            unsafe { expect_ok(ffi::kp_tokenizer_create(self.dict, &mut t), "kp_tokenizer_create") };
            t
        });
        let out = f(handle);
        self.pool.lock().unwrap().push(handle);
        out
    }
}

pub struct Tokenizer {
    pub dict: Dict,
    device: Arc<Device>,
}

/// Packs `inputs` the way the ABI takes a batch: one text buffer + `n + 1` byte offsets.
fn pack(inputs: &[&str]) -> (Vec<u8>, Vec<u64>) {
    let mut text = Vec::with_capacity(inputs.iter().map(|s| s.len()).sum());
    let mut offsets = Vec::with_capacity(inputs.len() + 1);
    offsets.push(0u64);
    for s in inputs {
        text.extend_from_slice(s.as_bytes());
        offsets.push(text.len() as u64);
    }
    (text, offsets)
}

/// `Vec<Token>` per sentence from the compact records (src/tokenizer.rs:22-43 builds the same fields from the
/// path nodes): a sentence's records are walked backwards from its EOS record -- position = sentence bytes,
/// start = the char count the EOS record carries -- subtracting every record's byte / char length; the
/// surface is the slice of the input the record covers (consecutive path nodes are adjacent, lattice.rs:124-125).
fn materialize(inputs: &[&str], r: &ffi::kp_result8) -> Vec<Vec<Token>> {
    let tok_off = unsafe { std::slice::from_raw_parts(r.tok_off, inputs.len() + 1) };
    let tokens = unsafe { std::slice::from_raw_parts(r.tokens, r.n_tokens as usize) };
    inputs
        .iter()
        .enumerate()
        .map(|(s, input)| {
            let recs = &tokens[tok_off[s] as usize..tok_off[s + 1] as usize];
            let mut out: Vec<Token> = Vec::with_capacity(recs.len());
            let (mut position, mut start) = (input.len(), 0usize);
            for (k, t) in recs.iter().enumerate().rev() {
                let id = (t.id_cls & 0x3FFF_FFFF) as isize;
                let token = if k == recs.len() - 1 {
                    // a non-empty path always ends with EOS (src/lattice.rs:144-153 starts from the last node)
                    start = t.byte_len as usize | (t.char_len as usize) << 16;
                    Token { id, class: TokenClass::Dummy, position, start, end: start + 3, surface: "EOS".to_string() }
                } else {
                    let end_byte = position;
                    position -= t.byte_len as usize;
                    start -= t.char_len as usize;
                    let class = if t.id_cls >> 30 == 1 { TokenClass::Known } else { TokenClass::Unknown };
                    Token {
                        id,
                        class,
                        position,
                        start,
                        end: start + t.char_len as usize,
                        surface: input[position..end_byte].to_string(),
                    }
                };
                out.push(token);
            }
            out.reverse();
            out
        })
        .collect()
}

impl Tokenizer {
    pub fn new(dict: Dict) -> Self {
        let device = device_of(&dict);
        Self { dict, device }
    }

    /// One line per call (src/bin/kanpyo.rs:115-122): the library serves this from its one-round-trip
    /// small-batch path (fused per-sentence kernel).
    pub fn tokenize(&self, input: &str) -> Vec<Token> {
        self.tokenize_batch(&[input]).pop().unwrap()
    }

    /// `tokenize` over many independent sentences in one device pass.
    pub fn tokenize_batch(&self, inputs: &[&str]) -> Vec<Vec<Token>> {
        let (text, offsets) = pack(inputs);
        self.device.with_handle(|handle| {
            let mut r = std::mem::MaybeUninit::<ffi::kp_result8>::uninit();
            let r = unsafe {
                expect_ok(
                    ffi::kp_tokenize_batch8(handle, text.as_ptr(), offsets.as_ptr(), inputs.len() as u64, r.as_mut_ptr()),
                    "kp_tokenize_batch8",
                );
                r.assume_init()
            };
            materialize(inputs, &r) // before the handle (which owns the result buffers) goes back to the pool
        })
    }

    /// The same batch split over every GPU of the box (byte-balanced contiguous sentence ranges; the
    /// dictionary reaches the other GPUs by one NCCL broadcast the first time this is called).
    pub fn tokenize_batch_all_gpus(&self, inputs: &[&str]) -> Vec<Vec<Token>> {
        let (text, offsets) = pack(inputs);
        let mut slot = self.device.shards.lock().unwrap();
        let group = *slot.get_or_insert_with(|| {
            let flat = Flat::of(&self.dict);
            let arrays = flat.arrays(&self.dict);
            let mut n: c_int = 0;
            let mut g = std::ptr::null_mut();
            unsafe {
                expect_ok(ffi::kp_device_count(&mut n), "kp_device_count");
                expect_ok(ffi::kp_shards_create(&arrays, std::ptr::null(), n, &mut g), "kp_shards_create");
            }
            g
        });
        let mut r = std::mem::MaybeUninit::<ffi::kp_result8>::uninit();
        let r = unsafe {
            expect_ok(
                ffi::kp_shards_tokenize(group, text.as_ptr(), offsets.as_ptr(), inputs.len() as u64, r.as_mut_ptr()),
                "kp_shards_tokenize",
            );
            r.assume_init()
        };
        materialize(inputs, &r)
    }
}
