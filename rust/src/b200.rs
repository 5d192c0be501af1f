//! b200.rs — `Icp2d` / `Icp3d` of tier4/icp_rust forwarded to libicp_b200.so (include/icp_b200.h, ABI 3).
//!
//! NOT compiled in the repository that ships it (no Rust toolchain there); written against
//! src/lib.rs:91-174, src/transform.rs:6-51 and src/types.rs:4-12 of the crate.
//! There is no CPU fallback: without a CUDA device `Icp2d::new` panics, as the header documents.

use core::ffi::{c_char, c_int, c_void};
use std::sync::{Mutex, OnceLock};

use crate::transform::Transform;
use crate::types::{Matrix2, Rotation2, Vector2, Vector3};

#[repr(C)] pub struct IcpB200Context { _p: [u8; 0] }
#[repr(C)] pub struct IcpB200Target { _p: [u8; 0] }

pub const ICP_B200_ABI_VERSION: c_int = 4;
pub const ICP_B200_OK: c_int = 0;
pub const ICP_B200_CORR_NONE: i32 = 0;
pub const ICP_B200_CORR_LAST: i32 = 1;
pub const ICP_B200_CORR_ALL: i32 = 2;
pub const ICP_B200_FLAG_ORDERED_SUMS: u32 = 1;
pub const ICP_B200_FLAG_SHARED_SOURCE: u32 = 2;
pub const ICP_B200_FLAG_NO_GUARD: u32 = 4;
pub const ICP_B200_FLAG_INDEPENDENT: u32 = 8;

/// icp_b200_batch (include/icp_b200.h): B independent `IcpD::new(dst_b).estimate(src_b, init_b, max_iter)`.
#[repr(C)]
pub struct IcpB200Batch {
    pub dim: u32, pub batch: u32, pub max_iter: u32, pub corr_mode: i32,
    pub max_src_points: u32, pub max_dst_points: u32,
    pub src: *const f64, pub src_off: *const u32,
    pub dst: *const f64, pub dst_off: *const u32,
    pub init_t: *const f64, pub out_t: *mut f64,
    pub out_inner: *mut u32, pub out_corr: *mut i32,
    pub out_corr_hash: *mut u64, pub out_status: *mut i32,
    pub flags: u32,
    pub out_cost: *mut f64,
}

/// icp_b200_call_info
#[repr(C)]
#[derive(Default, Clone, Copy)]
pub struct IcpB200CallInfo {
    pub kernel_launches: u32, pub kernel_ms: f32, pub h2d_bytes: u64, pub d2h_bytes: u64,
    pub grid_ctas: u32, pub block_threads: u32, pub smem_bytes: u32, pub redo_pairs: u32,
}

extern "C" {
    pub fn icp_b200_abi_version() -> c_int;
    pub fn icp_b200_status_string(status: c_int) -> *const c_char;
    pub fn icp_b200_device_count(count: *mut c_int) -> c_int;
    pub fn icp_b200_create(device: c_int, out: *mut *mut IcpB200Context) -> c_int;
    pub fn icp_b200_destroy(ctx: *mut IcpB200Context);
    pub fn icp_b200_last_call_info(ctx: *const IcpB200Context, info: *mut IcpB200CallInfo) -> c_int;
    pub fn icp_b200_last_error(ctx: *const IcpB200Context) -> *const c_char;
    // Icp2d::new(dst).estimate(src, init, max_iter) for every pair of a batch (src/lib.rs:97-130, 139-173)
    pub fn icp_b200_estimate_batch(ctx: *mut IcpB200Context, batch: *const IcpB200Batch) -> c_int;
    pub fn icp_b200_estimate_batch_device(ctx: *mut IcpB200Context, batch: *const IcpB200Batch,
                                          stream: *mut c_void, synchronize: c_int) -> c_int;
    // single pair, the reference's argument order (src/lib.rs:97, 105-110; 139, 148-153)
    pub fn icp_b200_icp2d_estimate(ctx: *mut IcpB200Context, dst_xy: *const f64, n_dst: u32,
                                   src_xy: *const f64, n_src: u32, initial_transform: *const f64,
                                   max_iter: u32, out_transform: *mut f64) -> c_int;
    pub fn icp_b200_icp3d_estimate(ctx: *mut IcpB200Context, dst_xyz: *const f64, n_dst: u32,
                                   src_xyz: *const f64, n_src: u32, initial_transform: *const f64,
                                   max_iter: u32, out_transform: *mut f64) -> c_int;
    // reusable target = the Icp2d<'a> / Icp3d<'a> value itself (src/lib.rs:91-102, 133-144)
    pub fn icp_b200_target_create(ctx: *mut IcpB200Context, dim: u32, dst: *const f64, n_dst: u32,
                                  out: *mut *mut IcpB200Target) -> c_int;
    pub fn icp_b200_target_destroy(target: *mut IcpB200Target);
    pub fn icp_b200_target_update(target: *mut IcpB200Target, add: *const f64, n_add: u32,
                                  remove_idx: *const u32, n_remove: u32) -> c_int;
    pub fn icp_b200_target_estimate_batch(target: *mut IcpB200Target, batch: *const IcpB200Batch) -> c_int;
    // the host-pointer call in two halves (one thread streaming batches through two contexts); the batch's
    // arrays must outlive the wait
    pub fn icp_b200_estimate_batch_submit(ctx: *mut IcpB200Context, batch: *const IcpB200Batch) -> c_int;
    pub fn icp_b200_estimate_batch_wait(ctx: *mut IcpB200Context) -> c_int;
    pub fn icp_b200_target_estimate_batch_submit(target: *mut IcpB200Target, batch: *const IcpB200Batch) -> c_int;
    pub fn icp_b200_target_estimate_batch_device(target: *mut IcpB200Target, batch: *const IcpB200Batch,
                                                 stream: *mut c_void, synchronize: c_int) -> c_int;
    // KdTree::search(&q).0 for every query (src/lib.rs:121,164)
    pub fn icp_b200_target_search(target: *mut IcpB200Target, queries: *const f64, n_queries: u64,
                                  out_index: *mut i32) -> c_int;
    pub fn icp_b200_target_search_device(target: *mut IcpB200Target, queries: *const f64, n_queries: u64,
                                         out_index: *mut i32, stream: *mut c_void, synchronize: c_int) -> c_int;
    // stats::calc_stddevs over ragged residual sets (src/stats.rs:49-60); test hook
    pub fn icp_b200_calc_stddevs_batch(ctx: *mut IcpB200Context, residuals_xy: *const f64, off: *const u32,
                                       batch: u32, out_sigma: *mut f64, out_median: *mut f64) -> c_int;
}

/// One context for the process (device 0): `Icp2d::new` is called once per frame by the examples
/// (examples/scan2d.rs:87) and must not create streams and events every time.  Calls on a context are
/// serialised (header "Threading"), hence the Mutex.
struct Ctx(*mut IcpB200Context);
unsafe impl Send for Ctx {}
static CONTEXT: OnceLock<Mutex<Ctx>> = OnceLock::new();

fn context() -> &'static Mutex<Ctx> {
    CONTEXT.get_or_init(|| {
        assert_eq!(unsafe { icp_b200_abi_version() }, ICP_B200_ABI_VERSION, "libicp_b200.so ABI mismatch");
        let mut ctx = core::ptr::null_mut();
        let rc = unsafe { icp_b200_create(0, &mut ctx) };
        assert_eq!(rc, ICP_B200_OK, "icp_b200_create failed ({rc}): there is no CPU fallback");
        Mutex::new(Ctx(ctx))
    })
}

fn to6(t: &Transform) -> [f64; 6] {
    let r = t.rot.matrix();
    [r[(0, 0)], r[(0, 1)], r[(1, 0)], r[(1, 1)], t.t[0], t.t[1]]
}

fn from6(a: &[f64; 6]) -> Transform {
    // same construction as so2::new_rotation2 (src/so2.rs:11-16) + Transform::from_rt (src/transform.rs:18-20)
    let rot = Rotation2::from_matrix_unchecked(Matrix2::new(a[0], a[1], a[2], a[3]));
    Transform::from_rt(&rot, &Vector2::new(a[4], a[5]))
}

fn check(rc: c_int) {
    // the reference panics on an empty target (src/lib.rs:122,165) and on NaN (src/stats.rs:12)
    assert_eq!(rc, ICP_B200_OK, "icp_b200 status {rc}");
}

pub struct Icp2d<'a> { dst: &'a [Vector2] }

impl<'a> Icp2d<'a> {
    pub fn new(dst: &'a [Vector2]) -> Self {                                       // src/lib.rs:97
        let _ = context();
        Icp2d { dst }
    }

    /// Estimates the transform that converts the `src` points to `dst`.           // src/lib.rs:104-110
    pub fn estimate(&self, src: &[Vector2], initial_transform: &Transform, max_iter: usize) -> Transform {
        let mut out = [0f64; 6];
        let ctx = context().lock().unwrap();
        // &[Vector2] is a contiguous [f64; 2] slice (nalgebra ArrayStorage, src/types.rs:4-12)
        check(unsafe {
            icp_b200_icp2d_estimate(ctx.0, self.dst.as_ptr() as *const f64, self.dst.len() as u32,
                                    src.as_ptr() as *const f64, src.len() as u32,
                                    to6(initial_transform).as_ptr(), max_iter as u32, out.as_mut_ptr())
        });
        from6(&out)
    }
}

pub struct Icp3d<'a> { dst: &'a [Vector3] }

impl<'a> Icp3d<'a> {
    pub fn new(dst: &'a [Vector3]) -> Self {                                       // src/lib.rs:139
        let _ = context();
        Icp3d { dst }
    }

    /// NN search in xyz, SE(2) solve on xy (src/lib.rs:148-173).
    pub fn estimate(&self, src: &[Vector3], initial_transform: &Transform, max_iter: usize) -> Transform {
        let mut out = [0f64; 6];
        let ctx = context().lock().unwrap();
        check(unsafe {
            icp_b200_icp3d_estimate(ctx.0, self.dst.as_ptr() as *const f64, self.dst.len() as u32,
                                    src.as_ptr() as *const f64, src.len() as u32,
                                    to6(initial_transform).as_ptr(), max_iter as u32, out.as_mut_ptr())
        });
        from6(&out)
    }
}

/// What the device is for: thousands of independent pairs in one call.  `ordered`: reference-order sums
/// (bit-identical transforms); default: guarded warp-tree sums (same decisions, transforms within 1e-9).
pub fn estimate_batch_2d(pairs: &[(&[Vector2], &[Vector2], Transform)], max_iter: usize, ordered: bool) -> Vec<Transform> {
    let b = pairs.len();
    let (mut src, mut dst) = (Vec::<f64>::new(), Vec::<f64>::new());
    let (mut so, mut dof) = (vec![0u32; b + 1], vec![0u32; b + 1]);
    let mut t0 = Vec::<f64>::with_capacity(6 * b);
    for (k, (s, d, init)) in pairs.iter().enumerate() {
        for p in s.iter() { src.extend_from_slice(&[p[0], p[1]]); }
        for p in d.iter() { dst.extend_from_slice(&[p[0], p[1]]); }
        so[k + 1] = so[k] + s.len() as u32;
        dof[k + 1] = dof[k] + d.len() as u32;
        t0.extend_from_slice(&to6(init));
    }
    let mut out = vec![0f64; 6 * b];
    let batch = IcpB200Batch {
        dim: 2, batch: b as u32, max_iter: max_iter as u32, corr_mode: ICP_B200_CORR_NONE,
        max_src_points: 0, max_dst_points: 0,
        src: src.as_ptr(), src_off: so.as_ptr(), dst: dst.as_ptr(), dst_off: dof.as_ptr(),
        init_t: t0.as_ptr(), out_t: out.as_mut_ptr(),
        out_inner: core::ptr::null_mut(), out_corr: core::ptr::null_mut(),
        out_corr_hash: core::ptr::null_mut(), out_status: core::ptr::null_mut(),
        flags: if ordered { ICP_B200_FLAG_ORDERED_SUMS } else { 0 },
        out_cost: core::ptr::null_mut(),
    };
    let ctx = context().lock().unwrap();
    check(unsafe { icp_b200_estimate_batch(ctx.0, &batch) });
    out.chunks_exact(6).map(|a| from6(a.try_into().unwrap())).collect()
}

/// Streaming batches from one thread: `push` enqueues a batch (icp_b200_estimate_batch_submit) on the next of
/// `DEPTH` contexts and returns the transforms of the batch pushed `DEPTH - 1` calls earlier, once it has been
/// waited for (icp_b200_estimate_batch_wait); `finish` drains.  The queue entry OWNS every array the batch points
/// to, so they outlive the wait as the header requires.  (For copies that overlap compute the Vecs should be
/// page-locked — cudaHostRegister on their buffers; left out of this sketch.)  DEPTH = 3: a batch that holds a
/// pair the knife-edge guard re-runs comes back about a batch time late, and two further batches in flight keep
/// the device busy meanwhile.
pub struct Pipeline {
    ctx: Vec<*mut IcpB200Context>,
    inflight: std::collections::VecDeque<InFlight>,
    next: usize,
    max_iter: u32,
}

struct InFlight {
    slot: usize,
    _src: Vec<f64>, _dst: Vec<f64>, _so: Vec<u32>, _dof: Vec<u32>, _t0: Vec<f64>,
    out: Vec<f64>,
}

impl Pipeline {
    pub const DEPTH: usize = 3;

    pub fn new(max_iter: usize) -> Self {
        let ctx = (0..Self::DEPTH).map(|_| {
            let mut c = core::ptr::null_mut();
            check(unsafe { icp_b200_create(0, &mut c) });
            c
        }).collect();
        Pipeline { ctx, inflight: Default::default(), next: 0, max_iter: max_iter as u32 }
    }

    fn retire(&mut self) -> Vec<Transform> {
        let f = self.inflight.pop_front().unwrap();
        check(unsafe { icp_b200_estimate_batch_wait(self.ctx[f.slot]) });
        f.out.chunks_exact(6).map(|a| from6(a.try_into().unwrap())).collect()
    }

    /// Enqueues `pairs`; returns the results of the oldest batch when DEPTH batches are in flight.
    pub fn push(&mut self, pairs: &[(&[Vector2], &[Vector2], Transform)]) -> Option<Vec<Transform>> {
        let b = pairs.len();
        let (mut src, mut dst) = (Vec::<f64>::new(), Vec::<f64>::new());
        let (mut so, mut dof) = (vec![0u32; b + 1], vec![0u32; b + 1]);
        let mut t0 = Vec::<f64>::with_capacity(6 * b);
        for (k, (s, d, init)) in pairs.iter().enumerate() {
            for p in s.iter() { src.extend_from_slice(&[p[0], p[1]]); }
            for p in d.iter() { dst.extend_from_slice(&[p[0], p[1]]); }
            so[k + 1] = so[k] + s.len() as u32;
            dof[k + 1] = dof[k] + d.len() as u32;
            t0.extend_from_slice(&to6(init));
        }
        let mut out = vec![0f64; 6 * b];
        let slot = self.next;
        self.next = (self.next + 1) % Self::DEPTH;
        let batch = IcpB200Batch {
            dim: 2, batch: b as u32, max_iter: self.max_iter, corr_mode: ICP_B200_CORR_NONE,
            max_src_points: 0, max_dst_points: 0,
            src: src.as_ptr(), src_off: so.as_ptr(), dst: dst.as_ptr(), dst_off: dof.as_ptr(),
            init_t: t0.as_ptr(), out_t: out.as_mut_ptr(),
            out_inner: core::ptr::null_mut(), out_corr: core::ptr::null_mut(),
            out_corr_hash: core::ptr::null_mut(), out_status: core::ptr::null_mut(),
            flags: 0, out_cost: core::ptr::null_mut(),
        };
        // the Vecs' heap buffers do not move when the Vecs are moved into the queue entry
        check(unsafe { icp_b200_estimate_batch_submit(self.ctx[slot], &batch) });
        self.inflight.push_back(InFlight { slot, _src: src, _dst: dst, _so: so, _dof: dof, _t0: t0, out });
        if self.inflight.len() == Self::DEPTH { Some(self.retire()) } else { None }
    }

    /// Waits for everything still in flight, oldest first.
    pub fn finish(&mut self) -> Vec<Vec<Transform>> {
        let mut all = Vec::new();
        while !self.inflight.is_empty() { all.push(self.retire()); }
        all
    }
}

impl Drop for Pipeline {
    fn drop(&mut self) {
        let _ = self.finish();
        for c in self.ctx.drain(..) { unsafe { icp_b200_destroy(c) }; }
    }
}
