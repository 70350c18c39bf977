//! Known-answer generator: runs THIS crate's own CPU path (no `cuda` feature needed) on the seeded
//! inputs of hodor_b200's `tests/golden/make_golden.py` and prints the same JSON schema as
//! `tests/golden/vectors.json`, so that one command pins hodor_b200's oracle, big-int model and CUDA
//! path against the real reference:
//!
//!     cargo test --release gen_vectors -- --nocapture --ignored > /tmp/out.txt
//!     sed -n '/^BEGIN_VECTORS$/,/^END_VECTORS$/p' /tmp/out.txt | sed '1d;$d' > vectors_ref.json
//!     python tests/golden/check_ref_vectors.py vectors_ref.json          # in the hodor_b200 repository
//!
//! It has to live inside the crate (`#[cfg(test)] mod gen_vectors;` in src/lib.rs, added by
//! hodor_cuda.patch) rather than under examples/: `Worker` sits in the `pub(crate)` module
//! `fft::multicore`, `bn256` is `pub(crate)` and `experiments` is private, so nothing outside the
//! crate can name them.
//!
//! Inputs: SplitMix64 stream, four limbs per element (limb 0 first), top limb masked to NUM_BITS,
//! rejection-sampled below the modulus and used DIRECTLY as the Montgomery representation
//! (`from_raw_repr`), exactly like `oracle/pymodel.py::random_mont_elements`.
//!
//! NOTE on cores: `(coset_)lde_using_multiple_cosets` seeds each worker chunk with
//! `coset_omega^chunk_index` (src/polynomials/mod.rs:448, :575), which is the right coset only when
//! the worker has at least `factor` cpus.  The generator therefore builds its Worker with
//! max(num_cpus, 16) cpus (`Worker::new_with_cpus`), the regime in which the reference's own
//! `test_coset_lde_correctness` passes.
#![cfg(test)]

use crypto::digest::Digest;
use crypto::sha2::Sha256;
use ff::{Field, PrimeField, PrimeFieldRepr};

use crate::domains::Domain;
use crate::fft::multicore::Worker;
use crate::fri::*;
use crate::iop::blake2s_trivial_iop::*;
use crate::iop::*;
use crate::polynomials::*;

mod bn254 {
    use ff::*;
    // pairing_ce's bn256::Fr declaration (this crate does not declare the field)
    #[derive(PrimeField)]
    #[PrimeFieldModulus = "21888242871839275222246405745257275088548364400416034343698204186575808495617"]
    #[PrimeFieldGenerator = "7"]
    pub struct Fr(FrRepr);
}

struct SplitMix64(u64);
impl SplitMix64 {
    fn next(&mut self) -> u64 {
        self.0 = self.0.wrapping_add(0x9E3779B97F4A7C15);
        let mut z = self.0;
        z = (z ^ (z >> 30)).wrapping_mul(0xBF58476D1CE4E5B9);
        z = (z ^ (z >> 27)).wrapping_mul(0x94D049BB133111EB);
        z ^ (z >> 31)
    }
}

fn random_mont_elements<F: PrimeField>(count: usize, seed: u64) -> Vec<F> {
    let mut gen = SplitMix64(seed);
    let modulus = F::char();
    let nlimbs = modulus.as_ref().len();
    assert!(nlimbs == 4, "the golden vectors cover the 4-limb fields");
    let mask = u64::max_value() >> (64 * nlimbs as u32 - F::NUM_BITS);
    let mut out = Vec::with_capacity(count);
    while out.len() < count {
        let mut repr = F::Repr::default();
        for limb in repr.as_mut().iter_mut() {
            *limb = gen.next();
        }
        repr.as_mut()[nlimbs - 1] &= mask;
        if repr < modulus {
            out.push(F::from_raw_repr(repr).expect("below the modulus"));
        }
    }
    out
}

fn mont_bytes<F: PrimeField>(values: &[F]) -> Vec<u8> {
    let mut out = Vec::with_capacity(values.len() * 32);
    for v in values.iter() {
        let mut buf = [0u8; 32];
        v.into_raw_repr().write_le(&mut buf[..]).expect("will write");
        out.extend_from_slice(&buf);
    }
    out
}

fn sha256_hex(bytes: &[u8]) -> String {
    let mut h = Sha256::new();
    h.input(bytes);
    h.result_str()
}

/// Python's hex(int) of the Montgomery representation
fn mont_hex<F: PrimeField>(v: &F) -> String {
    let repr = v.into_raw_repr();
    let mut s = String::new();
    for limb in repr.as_ref().iter().rev() {
        s.push_str(&format!("{:016x}", limb));
    }
    let t = s.trim_start_matches('0');
    format!("0x{}", if t.is_empty() { "0" } else { t })
}

fn quote_list(items: &[String]) -> String {
    let q: Vec<String> = items.iter().map(|s| format!("\"{}\"", s)).collect();
    format!("[{}]", q.join(", "))
}

fn cases_for_field<F: PrimeField>(field_id: usize, worker: &Worker, out: &mut Vec<String>) {
    // ---- ntt: best_fft with the domain generator (src/fft/fft.rs:5-66) ----------------------
    for (ci, ln) in [0u32, 1, 2, 5, 8, 11, 12, 13].iter().enumerate() {
        let seed = 0x3DBE62598D313D76u64 + ci as u64;
        let mut a = random_mont_elements::<F>(1 << ln, seed);
        let omega = Domain::<F>::new_for_size(1u64 << ln).expect("domain").generator;
        crate::fft::best_fft(&mut a, worker, &omega, *ln, None);
        out.push(format!(
            "{{\"kind\": \"ntt\", \"field\": {}, \"log_n\": {}, \"seed\": {}, \"sha256\": \"{}\", \"first\": \"{}\", \"last\": \"{}\"}}",
            field_id, ln, seed, sha256_hex(&mont_bytes(&a)), mont_hex(&a[0]), mont_hex(&a[a.len() - 1])
        ));
    }
    // ---- (coset) LDE (src/polynomials/mod.rs:343-352) ----------------------------------------
    for (ci, (ln, factor, coset)) in [(3u32, 2usize, false), (4, 8, true), (6, 8, true), (9, 8, true), (5, 16, false)].iter().enumerate() {
        let seed = 0x1000u64 + ci as u64;
        let a = random_mont_elements::<F>(1 << ln, seed);
        let poly = Polynomial::<F, Coefficients>::from_coeffs(a).expect("poly");
        let lde = if *coset { poly.coset_lde(worker, *factor) } else { poly.lde(worker, *factor) }.expect("lde");
        out.push(format!(
            "{{\"kind\": \"lde\", \"field\": {}, \"log_n\": {}, \"factor\": {}, \"coset\": {}, \"seed\": {}, \"sha256\": \"{}\"}}",
            field_id, ln, factor, coset, seed, sha256_hex(&mont_bytes(lde.as_ref()))
        ));
    }
    // ---- Blake2sIopTree::create (src/iop/blake2s_trivial_iop.rs:131-234) -----------------------
    for (ci, ln) in [1u32, 2, 5, 9, 13].iter().enumerate() {
        let seed = 0x2000u64 + ci as u64;
        let a = random_mont_elements::<F>(1 << ln, seed);
        let tree = Blake2sIopTree::<F>::create(&a);
        let mut all = Vec::with_capacity(32 << ln);
        for n in tree.nodes().iter() {
            all.extend_from_slice(n);
        }
        out.push(format!(
            "{{\"kind\": \"merkle\", \"field\": {}, \"log_n\": {}, \"seed\": {}, \"root\": \"{}\", \"nodes_sha256\": \"{}\", \"challenge\": \"{}\"}}",
            field_id, ln, seed, hex::encode(&tree.get_root()[..]), sha256_hex(&all), mont_hex(&tree.get_challenge_scalar_from_root())
        ));
    }
    // ---- NaiveFriIop::proof_from_lde_by_values (src/fri/fri_on_values.rs:11-159) ---------------
    for (ci, (ln, factor, oc)) in [(4u32, 4usize, 2usize), (8, 8, 1), (10, 16, 4)].iter().enumerate() {
        let seed = 0x3000u64 + ci as u64;
        let a = random_mont_elements::<F>(1 << ln, seed);
        let values = Polynomial::<F, Values>::from_values(a).expect("poly");
        let proto = NaiveFriIop::<F, TrivialBlake2sIOP<F>>::proof_from_lde_by_values(&values, *factor, *oc, worker).expect("fri");
        let roots: Vec<String> = proto.get_roots().iter().map(|r| hex::encode(&r[..])).collect();
        let challenges: Vec<String> = proto.challenges.iter().map(|c| mont_hex(c)).collect();
        let finals: Vec<String> = proto.final_coefficients.iter().map(|c| mont_hex(c)).collect();
        let vals: Vec<String> = proto.intermediate_values.iter().map(|v| sha256_hex(&mont_bytes(v.as_ref()))).collect();
        out.push(format!(
            "{{\"kind\": \"fri\", \"field\": {}, \"log_n\": {}, \"lde_factor\": {}, \"out_coeffs\": {}, \"seed\": {}, \"roots\": {}, \
             \"challenges\": {}, \"final_root\": \"{}\", \"final_coefficients\": {}, \"values_sha256\": {}}}",
            field_id, ln, factor, oc, seed, quote_list(&roots), quote_list(&challenges), hex::encode(&proto.final_root[..]),
            quote_list(&finals), quote_list(&vals)
        ));
    }
    // ---- batch_inversion / evaluate_at (src/polynomials/mod.rs:889-954, 685-711) ----------------
    for (ci, ln) in [0u32, 3, 8, 12].iter().enumerate() {
        let seed = 0x4000u64 + ci as u64;
        let a = random_mont_elements::<F>(1 << ln, seed);
        let mut inv = Polynomial::<F, Values>::from_values(a.clone()).expect("poly");
        inv.batch_inversion(worker).expect("no zero in a random vector");
        out.push(format!(
            "{{\"kind\": \"batch_inversion\", \"field\": {}, \"log_n\": {}, \"seed\": {}, \"sha256\": \"{}\", \"first\": \"{}\"}}",
            field_id, ln, seed, sha256_hex(&mont_bytes(inv.as_ref())), mont_hex(&inv.as_ref()[0])
        ));
        let z = random_mont_elements::<F>(1, seed + 0x100)[0];
        let value = Polynomial::<F, Coefficients>::from_coeffs(a).expect("poly").evaluate_at(worker, z);
        out.push(format!(
            "{{\"kind\": \"evaluate_at\", \"field\": {}, \"log_n\": {}, \"seed\": {}, \"point_seed\": {}, \"value\": \"{}\"}}",
            field_id, ln, seed, seed + 0x100, mont_hex(&value)
        ));
    }
}

#[test]
#[ignore] // a generator, not a check: run it explicitly (see the module comment)
fn gen_vectors() {
    let cpus = std::cmp::max(Worker::new().num_cpus() as usize, 16);
    let worker = Worker::new_with_cpus(cpus);
    let mut cases = vec![];
    cases_for_field::<crate::bn256::Fr>(0, &worker, &mut cases);
    cases_for_field::<bn254::Fr>(1, &worker, &mut cases);
    cases_for_field::<crate::experiments::Fr>(2, &worker, &mut cases);
    println!("BEGIN_VECTORS");
    println!("{{\"generator\": \"matter-labs/hodor CPU path (cargo test gen_vectors), splitmix64 seeds of tests/golden/make_golden.py\",");
    println!(" \"cases\": [");
    for (i, c) in cases.iter().enumerate() {
        println!("  {}{}", c, if i + 1 < cases.len() { "," } else { "" });
    }
    println!(" ]}}");
    println!("END_VECTORS");
}
