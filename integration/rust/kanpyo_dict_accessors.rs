//! Read-only accessors a maintainer adds to `kanpyo-dict` so that the `kanpyo` crate can hand the
//! dictionary's arrays to the CUDA library (the fields are private today).  SOURCE ONLY.

// kanpyo-dict/src/trie/da.rs (DoubleArray(Vec<Node>), :20)
impl DoubleArray {
    pub fn nodes(&self) -> &[Node] {
        &self.0
    }
}

// kanpyo-dict/src/index.rs (IndexTable { da, dup }, :10-13)
impl IndexTable {
    pub fn da_nodes(&self) -> &[crate::trie::da::Node] {
        self.da.nodes()
    }
    pub fn dup_map(&self) -> &std::collections::BTreeMap<KeywordID, usize> {
        &self.dup
    }
}

// kanpyo-dict/src/morph.rs (Morphs(Vec<Morph>), :24)
impl Morphs {
    pub fn as_slice(&self) -> &[Morph] {
        &self.0
    }
}

// kanpyo-dict/src/connection.rs (ConnectionTable { row, col, data }, :5-9)
impl ConnectionTable {
    pub fn shape(&self) -> (usize, usize) {
        (self.row, self.col)
    }
    pub fn as_slice(&self) -> &[i16] {
        &self.data
    }
}
