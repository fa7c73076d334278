//! GPU drop-in for `cdivsufsort` / `divsufsort` (source only).
//!
//! `sort` / `sort_in_place` have the signatures of crates/cdivsufsort/src/lib.rs:9-30 and
//! crates/divsufsort/src/lib.rs:20-29.  `GpuSuffixArray` implements `sacabase::StringIndex`
//! (crates/sacabase/src/lib.rs:160-163) on a device-resident index, and
//! `GpuPartitionedSuffixArray` mirrors `sacapart::PartitionedSuffixArray`
//! (crates/sacapart/src/lib.rs:26-98).
use sacabase::{LongestCommonSubstring, StringIndex, SuffixArray};
use std::os::raw::c_char;

#[repr(C)]
pub struct GsaIndex {
    _private: [u8; 0],
}
#[repr(C)]
pub struct GsaPart {
    _private: [u8; 0],
}

extern "C" {
    fn gsa_divsufsort(T: *const u8, SA: *mut i32, n: i32) -> i32;
    fn gsa_index_from_parts(T: *const u8, n: i64, SA: *const i32, sa_len: i64, device: i32, out: *mut *mut GsaIndex) -> i32;
    fn gsa_index_verify(ix: *const GsaIndex, bad_index: *mut i64) -> i32;
    fn gsa_index_destroy(ix: *mut GsaIndex);
    fn gsa_lsm_batch(ix: *const GsaIndex, pats: *const u8, pat_off: *const u64, q: u64, out_start: *mut u64, out_len: *mut u32) -> i32;
    fn gsa_search_all_batch(ix: *const GsaIndex, pats: *const u8, pat_off: *const u64, q: u64, out_left: *mut i32, out_count: *mut i32) -> i32;
    fn gsa_part_create(T: *const u8, n: u64, num_partitions: u64, devices: *const i32, ndev: i32, out: *mut *mut GsaPart) -> i32;
    fn gsa_part_num_partitions(p: *const GsaPart) -> u64;
    fn gsa_part_lsm_batch(p: *mut GsaPart, pats: *const u8, pat_off: *const u64, q: u64, out_start: *mut u64, out_len: *mut u32) -> i32;
    fn gsa_part_destroy(p: *mut GsaPart);
    fn gsa_divbwt(T: *const u8, U: *mut u8, A: *mut i32, n: i32) -> i32;
    fn gsa_inverse_bw_transform(T: *const u8, U: *mut u8, A: *mut i32, n: i32, idx: i32) -> i32;
    fn gsa_lcp(T: *const u8, SA: *const i32, LCP: *mut i32, n: i32, device: i32) -> i32;
    fn gsa_last_error() -> *const c_char;
}

fn last_error() -> String {
    unsafe { std::ffi::CStr::from_ptr(gsa_last_error()).to_string_lossy().into_owned() }
}

/// Burrows-Wheeler transform with libdivsufsort's `divbwt` convention (divsufsort.c:372-405):
/// returns the transformed string and the primary index.
pub fn bwt(text: &[u8]) -> (Vec<u8>, i32) {
    assert!(text.len() < i32::max_value() as usize);
    let mut u = vec![0u8; text.len()];
    let ret = unsafe { gsa_divbwt(text.as_ptr(), u.as_mut_ptr(), std::ptr::null_mut(), text.len() as i32) };
    assert!(ret >= 0, "{}", last_error());
    (u, ret)
}

/// Inverse of [`bwt`] (`inverse_bw_transform`, utils.c:111-156).
pub fn inverse_bwt(u: &[u8], primary_index: i32) -> Vec<u8> {
    let mut t = vec![0u8; u.len()];
    let ret = unsafe { gsa_inverse_bw_transform(u.as_ptr(), t.as_mut_ptr(), std::ptr::null_mut(), u.len() as i32, primary_index) };
    assert_eq!(0, ret, "{}", last_error());
    t
}

/// LCP array of a suffix array: `lcp[0] = 0`, `lcp[j]` = common prefix length of the suffixes
/// at `sa[j-1]` and `sa[j]`.
pub fn lcp(text: &[u8], sa: &[i32]) -> Vec<i32> {
    assert_eq!(text.len(), sa.len(), "text and suffix array should have same len");
    let mut out = vec![0i32; sa.len()];
    if !sa.is_empty() {
        let ret = unsafe { gsa_lcp(text.as_ptr(), sa.as_ptr(), out.as_mut_ptr(), sa.len() as i32, 0) };
        assert_eq!(0, ret, "{}", last_error());
    }
    out
}

/// Sort suffixes of `text` and store their lexographic order in the given suffix array `sa`.
/// Will panic if `sa.len()` != `text.len()`  (cdivsufsort lib.rs:9-23)
pub fn sort_in_place(text: &[u8], sa: &mut [i32]) {
    assert_eq!(text.len(), sa.len(), "text and suffix array should have same len");
    assert!(
        text.len() < i32::max_value() as usize,
        "text too large, should not exceed {} bytes",
        i32::max_value() - 1
    );
    let ret = unsafe { gsa_divsufsort(text.as_ptr(), sa.as_mut_ptr(), text.len() as i32) };
    assert_eq!(0, ret, "{}", last_error());
}

/// Sort suffixes (cdivsufsort lib.rs:26-30)
pub fn sort<'a>(text: &'a [u8]) -> SuffixArray<'a, i32> {
    let mut sa = vec![0; text.len()];
    sort_in_place(text, &mut sa);
    SuffixArray::new(text, sa)
}

/// A suffix array whose text and entries also live in GPU memory, for batched queries.
pub struct GpuSuffixArray<'a> {
    text: &'a [u8],
    sa: Vec<i32>,
    ix: *mut GsaIndex,
}

impl<'a> GpuSuffixArray<'a> {
    pub fn new(text: &'a [u8], sa: Vec<i32>, device: i32) -> Self {
        assert_eq!(text.len(), sa.len(), "text and suffix array should have same len");
        let mut ix = std::ptr::null_mut();
        // the library re-checks the length and range-checks every entry on the device (GSA_EPANIC)
        let rc = unsafe { gsa_index_from_parts(text.as_ptr(), text.len() as i64, sa.as_ptr(), sa.len() as i64, device, &mut ix) };
        assert_eq!(0, rc, "{}", last_error());
        Self { text, sa, ix }
    }
    pub fn sort(text: &'a [u8], device: i32) -> Self {
        let (text, sa) = sort(text).into_parts();
        Self::new(text, sa, device)
    }
    pub fn into_parts(mut self) -> (&'a [u8], Vec<i32>) {
        (self.text, std::mem::replace(&mut self.sa, Vec::new()))
    }
    pub fn text(&self) -> &[u8] {
        self.text
    }
    /// O(n) check on the GPU; Err(i) = slot where suf(SA(i)) < suf(SA(i+1)) fails.
    pub fn verify(&self) -> Result<(), usize> {
        let mut bad = -1i64;
        match unsafe { gsa_index_verify(self.ix, &mut bad) } {
            0 => Ok(()),
            1 => Err(bad as usize),
            rc => panic!("gsa_index_verify: {} {}", rc, last_error()),
        }
    }
    /// One (start, len) per needle; needle q is `pats[off[q]..off[q + 1]]`.
    pub fn longest_substring_match_batch(&self, pats: &[u8], off: &[u64]) -> (Vec<u64>, Vec<u32>) {
        let q = off.len() - 1;
        let (mut start, mut len) = (vec![0u64; q], vec![0u32; q]);
        let rc = unsafe { gsa_lsm_batch(self.ix, pats.as_ptr(), off.as_ptr(), q as u64, start.as_mut_ptr(), len.as_mut_ptr()) };
        assert_eq!(0, rc, "{}", last_error());
        (start, len)
    }
    /// libdivsufsort `sa_search` semantics: occurrences of pattern q are `sa[left..left + count]`.
    pub fn search_all_batch(&self, pats: &[u8], off: &[u64]) -> (Vec<i32>, Vec<i32>) {
        let q = off.len() - 1;
        let (mut left, mut count) = (vec![0i32; q], vec![0i32; q]);
        let rc = unsafe { gsa_search_all_batch(self.ix, pats.as_ptr(), off.as_ptr(), q as u64, left.as_mut_ptr(), count.as_mut_ptr()) };
        assert_eq!(0, rc, "{}", last_error());
        (left, count)
    }
    pub fn search_all(&self, pattern: &[u8]) -> &[i32] {
        let (left, count) = self.search_all_batch(pattern, &[0, pattern.len() as u64]);
        if count[0] <= 0 {
            &self.sa[..0]
        } else {
            &self.sa[left[0] as usize..(left[0] + count[0]) as usize]
        }
    }
    pub fn contains(&self, pattern: &[u8]) -> bool {
        self.search_all_batch(pattern, &[0, pattern.len() as u64]).1[0] > 0
    }
}

impl<'a> StringIndex<'a> for GpuSuffixArray<'a> {
    fn longest_substring_match(&self, needle: &[u8]) -> LongestCommonSubstring<'a> {
        let (start, len) = self.longest_substring_match_batch(needle, &[0, needle.len() as u64]);
        LongestCommonSubstring { text: self.text, start: start[0] as usize, len: len[0] as usize }
    }
}

impl<'a> Drop for GpuSuffixArray<'a> {
    fn drop(&mut self) {
        unsafe { gsa_index_destroy(self.ix) }
    }
}

/// `sacapart::PartitionedSuffixArray` with the shards resident on `devices` (shard i on
/// `devices[i % devices.len()]`).
pub struct GpuPartitionedSuffixArray<'a> {
    text: &'a [u8],
    h: *mut GsaPart,
}

impl<'a> GpuPartitionedSuffixArray<'a> {
    pub fn new(text: &'a [u8], num_partitions: usize, devices: &[i32]) -> Self {
        assert!(num_partitions != 0, "attempt to divide by zero"); // sacapart lib.rs:43
        let mut h = std::ptr::null_mut();
        let rc = unsafe {
            gsa_part_create(text.as_ptr(), text.len() as u64, num_partitions as u64, devices.as_ptr(), devices.len() as i32, &mut h)
        };
        assert_eq!(0, rc, "{}", last_error());
        Self { text, h }
    }
    pub fn num_partitions(&self) -> usize {
        unsafe { gsa_part_num_partitions(self.h) as usize }
    }
    pub fn longest_substring_match_batch(&self, pats: &[u8], off: &[u64]) -> (Vec<u64>, Vec<u32>) {
        let q = off.len() - 1;
        let (mut start, mut len) = (vec![0u64; q], vec![0u32; q]);
        let rc = unsafe { gsa_part_lsm_batch(self.h, pats.as_ptr(), off.as_ptr(), q as u64, start.as_mut_ptr(), len.as_mut_ptr()) };
        assert_eq!(0, rc, "partitioned suffix arrays should always find at least one longest common substring");
        (start, len)
    }
}

impl<'a> StringIndex<'a> for GpuPartitionedSuffixArray<'a> {
    fn longest_substring_match(&self, needle: &[u8]) -> LongestCommonSubstring<'a> {
        let (start, len) = self.longest_substring_match_batch(needle, &[0, needle.len() as u64]);
        LongestCommonSubstring { text: self.text, start: start[0] as usize, len: len[0] as usize }
    }
}

impl<'a> Drop for GpuPartitionedSuffixArray<'a> {
    fn drop(&mut self) {
        unsafe { gsa_part_destroy(self.h) }
    }
}
