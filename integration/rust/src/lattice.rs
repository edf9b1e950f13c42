//! Drop-in for the public surface of the reference's `src/lattice.rs` that callers use
//! (`Lattice::build`, `pub nodes`, `pub edges`, `Lattice::viterbi`; src/bin/kanpyo.rs:146,
//! src/graphviz.rs:31-35), served by `kp_lattice_dump`.  SOURCE ONLY.
use kanpyo_dict::{dict::Dict, morph::Morph};

use crate::ffi;
use crate::lattice::node::{Node, Word};
use crate::tokenizer::device_of;

pub mod node; // unchanged: src/lattice/node.rs

pub struct Lattice<'a> {
    pub dict: &'a Dict,
    pub nodes: Vec<Node>,
    pub edges: Vec<Vec<usize>>,
    pre: Vec<Option<usize>>, // pre_nodes of viterbi(), computed on the device
}

impl<'a> Lattice<'a> {
    /// The reference's signature (src/lattice.rs:101), so `src/bin/kanpyo.rs:146` and `src/graphviz.rs` compile
    /// unchanged: the device state of `dict` is found (or created) through the registry in `tokenizer.rs`, and a
    /// tokenizer handle is borrowed from its pool for the duration of the call.
    pub fn build(dict: &'a Dict, input: &str) -> Self {
        let device = device_of(dict);
        device.with_handle(|handle| Self::build_with(handle, dict, input))
    }

    /// `handle`: a `kp_tokenizer*` nobody else uses during the call.
    pub fn build_with(handle: *mut ffi::kp_tokenizer, dict: &'a Dict, input: &str) -> Self {
        let mut la = std::mem::MaybeUninit::<ffi::kp_lattice>::uninit();
        let la = unsafe {
            let rc = ffi::kp_lattice_dump(handle, input.as_ptr(), input.len() as u64, la.as_mut_ptr());
            assert_eq!(rc, ffi::KP_OK, "kp_lattice_dump failed");
            la.assume_init()
        };
        // the node table is copied out below, before the handle returns to the pool
        let raw = unsafe { std::slice::from_raw_parts(la.nodes, la.n_nodes as usize) };
        let n_chars = input.chars().count();
        let byte_of_char: Vec<usize> =
            input.char_indices().map(|(b, _)| b).chain(std::iter::once(input.len())).collect();
        let mut edges = vec![vec![]; n_chars + 2];
        let mut nodes = Vec::with_capacity(raw.len());
        let mut pre = Vec::with_capacity(raw.len());
        for (i, n) in raw.iter().enumerate() {
            let morph = Morph::new(n.left_id, n.right_id, n.cost);
            let (byte_pos, char_pos) = (n.byte_pos as usize, n.char_pos as usize);
            nodes.push(match n.cls {
                0 => Node::Dummy { byte_pos, char_pos, morph },
                cls => {
                    let end_byte = byte_of_char[n.end_char as usize];
                    let word = Word {
                        id: n.id as isize,
                        byte_pos,
                        char_pos,
                        morph,
                        surface: input[byte_pos..end_byte].to_string(),
                    };
                    if cls == 1 { Node::Known(word) } else { Node::Unknown(word) }
                }
            });
            edges[n.end_char as usize].push(i);
            pre.push(if n.pre < 0 { None } else { Some(n.pre as usize) });
        }
        Self { dict, nodes, edges, pre }
    }

    /// Best path without BOS (src/lattice.rs:144-153).
    pub fn viterbi(&self) -> Vec<Node> {
        let mut path = vec![];
        let mut pos = self.nodes.len() - 1;
        while let Some(p) = self.pre[pos] {
            path.push(self.nodes[pos].clone());
            pos = p;
        }
        path.reverse();
        path
    }
}
