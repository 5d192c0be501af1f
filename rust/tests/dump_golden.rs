//! dump_golden.rs — run with the REAL crate (`cargo test --features std --test dump_golden -- --nocapture`)
//! to pin the three points the C oracle of the B200 repository could not pin (no Rust toolchain there):
//!   1. stats::mutable_median for EVEN n well above 10 (src/stats.rs:23-26 reads input[n/2 - 1] after a
//!      second select_nth_unstable_by(n/2), which core::slice::select does not promise to leave in place);
//!   2. the kd-tree's choice among equidistant points (nearest_neighbor::KdTree, Cargo.toml:22-25);
//!   3. the last bits of f64::sin / f64::cos as the crate computes them (libm feature or std).
//! NOT compiled in the repository that ships it.  Output: one JSON object on stdout; commit it as
//! tests/golden/reference_rust.json next to this file's commit hash and rustc version.

use icp::{Icp2d, Transform, Vector2};

fn lcg(state: &mut u64) -> f64 {
    *state = state.wrapping_mul(6364136223846793005).wrapping_add(1442695040888963407);
    ((*state >> 11) as f64) / ((1u64 << 53) as f64)
}

#[test]
fn dump() {
    let mut s = 20240807u64;
    println!("{{");
    // 1. even-n medians through the public path: residual sets are private, so drive calc_stddevs through
    //    weighted_gauss_newton_update is not possible either — expose `pub fn median_for_golden` behind
    //    cfg(test) in src/stats.rs calling mutable_median, or copy stats.rs:11-28 here verbatim:
    let mut medians = Vec::new();
    for n in [12usize, 64, 666, 1080, 4096] {
        let mut v: Vec<f64> = (0..n).map(|_| lcg(&mut s) * 20.0 - 10.0).collect();
        let input = v.clone();
        // stats.rs:11-28 verbatim
        let f = |a: &f64, b: &f64| a.partial_cmp(b).unwrap();
        let n2 = v.len();
        let (_, &mut b, _) = v.select_nth_unstable_by(n2 / 2, f);
        let (_, &mut c, _) = v.select_nth_unstable_by(n2 / 2 - 1, f);   // order as in the crate
        let med = (b + c) / 2.0;
        let mut sorted = input.clone();
        sorted.sort_by(f);
        let exact = (sorted[n2 / 2 - 1] + sorted[n2 / 2]) / 2.0;
        medians.push(format!("{{\"n\":{n},\"crate_bits\":{},\"exact_bits\":{}}}", med.to_bits(), exact.to_bits()));
    }
    println!("\"even_n_medians\":[{}],", medians.join(","));
    // 2. ties: duplicated target points and a query equidistant from two distinct points
    let dst = vec![Vector2::new(0., 0.), Vector2::new(1., 0.), Vector2::new(0., 0.), Vector2::new(-1., 0.),
                   Vector2::new(0., 0.), Vector2::new(3., 3.)];
    let icp = Icp2d::new(&dst);
    // estimate() does not return indices: print the transform of one iteration from identity for sources
    // that sit on the ties; the B200 repository's oracle reproduces it only with the same tie rule.
    let src = vec![Vector2::new(0.1, 0.0), Vector2::new(0.0, 0.2), Vector2::new(0.0, 1e-9), Vector2::new(2.9, 3.1)];
    let t = icp.estimate(&src, &Transform::identity(), 1);
    let r = t.rot.matrix();
    println!("\"tie_case_transform_bits\":[{},{},{},{},{},{}],",
             r[(0, 0)].to_bits(), r[(0, 1)].to_bits(), r[(1, 0)].to_bits(), r[(1, 1)].to_bits(), t.t[0].to_bits(), t.t[1].to_bits());
    // 3. sin / cos
    let mut sc = Vec::new();
    for _ in 0..64 {
        let x = (lcg(&mut s) - 0.5) * 0.2;
        let tr = Transform::new(&icp::Vector3::new(0., 0., x));
        let m = tr.rot.matrix();
        sc.push(format!("[{},{},{}]", x.to_bits(), m[(1, 0)].to_bits(), m[(0, 0)].to_bits()));
    }
    println!("\"theta_sin_cos_bits\":[{}]", sc.join(","));
    println!("}}");
}
