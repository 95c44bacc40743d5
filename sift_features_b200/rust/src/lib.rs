//! Drop-in for the extraction path of `sift-features` on NVIDIA B200: the same public items as the
//! crate's `src/lib.rs:39-81,124-177,785` (`sift`, `sift_with_processing`, `precompute_images`,
//! `sift_with_precomputed`, `compute_descriptor`, `SiftResult`, `KeyPoint`), implemented by calls into
//! `libsift_b200.so` (C ABI: include/sift_b200.h).  NOT compiled in the build image (no rustc); kept as
//! the reference-side binding described in INTEGRATION.md.
use image::GrayImage;
use ndarray::{Array2, Array3, ArrayView2};
use std::os::raw::{c_char, c_int};

#[repr(C)]
#[derive(Debug, Clone, Copy, PartialEq, PartialOrd)]
pub struct KeyPoint {
    pub x: f32,
    pub y: f32,
    pub size: f32,
    pub angle: f32,
    pub response: f32,
}

#[derive(Debug, Clone, PartialEq)]
pub struct SiftResult {
    pub keypoints: Vec<KeyPoint>,
    /// `(keypoints.len(), 128)`, same order as `keypoints`
    pub descriptors: Array2<u8>,
}

#[repr(C)]
struct Sb200Ctx {
    _private: [u8; 0],
}
#[repr(C)]
struct Sb200Result {
    n: u64,
    n_images: u32,
    offsets: *const u64,
    keypoints: *const KeyPoint,
    descriptors: *const u8,
}
#[repr(C)]
struct Sb200DescIn {
    x: f32,
    y: f32,
    scale: f32,
    orientation: f32,
}

extern "C" {
    fn sb200_create(device: c_int, max_w: u32, max_h: u32, max_batch: u32, max_kp: u32, out: *mut *mut Sb200Ctx) -> c_int;
    fn sb200_destroy(ctx: *mut Sb200Ctx);
    fn sb200_last_error(ctx: *const Sb200Ctx) -> *const c_char;
    fn sb200_set_processing(ctx: *mut Sb200Ctx, processing: c_int) -> c_int;
    fn sb200_pyramid_info(ctx: *mut Sb200Ctx, n_octaves: *mut u32, widths: *mut u32, heights: *mut u32, cap: u32) -> c_int;
    fn sb200_pyramid_layer(ctx: *mut Sb200Ctx, octave: u32, layer: u32, out: *mut f32) -> c_int;
    fn sb200_pyramid_dog(ctx: *mut Sb200Ctx, octave: u32, layer: u32, out: *mut f32) -> c_int;
    fn sb200_extract(ctx: *mut Sb200Ctx, gray: *const u8, w: u32, h: u32, stride: u32, limit: i64, out: *mut Sb200Result) -> c_int;
    fn sb200_precompute(ctx: *mut Sb200Ctx, gray: *const u8, w: u32, h: u32, stride: u32) -> c_int;
    fn sb200_extract_precomputed(ctx: *mut Sb200Ctx, limit: i64, out: *mut Sb200Result) -> c_int;
    fn sb200_compute_descriptors(ctx: *mut Sb200Ctx, img: *const f32, w: u32, h: u32, stride: u32,
                                 kps: *const Sb200DescIn, n: u64, out: *mut u8) -> c_int;
    fn sb200_jpeg_info(ctx: *mut Sb200Ctx, jpeg: *const u8, length: u64, w: *mut u32, h: *mut u32, components: *mut u32) -> c_int;
    fn sb200_extract_batch_jpeg(ctx: *mut Sb200Ctx, jpegs: *const *const u8, lengths: *const u64, n: u32, limit: i64,
                                out: *mut Sb200Result) -> c_int;
    fn sb200_match_descriptors(ctx: *mut Sb200Ctx, query: *const u8, n_query: u64, train: *const u8, n_train: u64,
                               out: *mut DMatch, cap: u64, n_out: *mut u64) -> c_int;
}

/// One mutual nearest-neighbour pair (the fields of `cv::DMatch` the examples use; `dist2` is the exact squared L2).
#[repr(C)]
#[derive(Debug, Clone, Copy, PartialEq, Eq)]
pub struct DMatch {
    pub query: u32,
    pub train: u32,
    pub dist2: u32,
}

/// `BFMatcher::new(NORM_L2, true)` + `match_(query, train)` of `examples/sift-match.rs:30-35` on the GPU.
pub fn match_descriptors(query: &Array2<u8>, train: &Array2<u8>) -> Vec<DMatch> {
    assert!(query.ncols() == 128 && train.ncols() == 128);
    let (q, t) = (query.as_standard_layout(), train.as_standard_layout());
    let ctx = Context::new(8, 8);
    let mut out = Vec::<DMatch>::with_capacity(q.nrows());
    let mut n = 0u64;
    let st = unsafe { sb200_match_descriptors(ctx.0, q.as_ptr(), q.nrows() as u64, t.as_ptr(), t.nrows() as u64,
                                              out.as_mut_ptr(), out.capacity() as u64, &mut n) };
    ctx.check(st);
    unsafe { out.set_len(n as usize) };
    out
}

/// `image::load_from_memory(bytes)?.grayscale()` + `sift()` (`src/lib.rs:1012-1019`, `examples/run-sift.rs:8-19`) with
/// the JPEG decode and the luma conversion on the device.
pub fn sift_jpeg(jpeg: &[u8], features_limit: Option<usize>) -> SiftResult {
    let probe = Context::new(8, 8);
    let (mut w, mut h, mut c) = (0u32, 0u32, 0u32);
    probe.check(unsafe { sb200_jpeg_info(probe.0, jpeg.as_ptr(), jpeg.len() as u64, &mut w, &mut h, &mut c) });
    let ctx = Context::new(w, h);
    let mut r = std::mem::MaybeUninit::<Sb200Result>::zeroed();
    let (ptr, len) = (jpeg.as_ptr(), jpeg.len() as u64);
    let st = unsafe { sb200_extract_batch_jpeg(ctx.0, &ptr, &len, 1, features_limit.map_or(-1, |l| l as i64), r.as_mut_ptr()) };
    ctx.check(st);
    Context::take(unsafe { &r.assume_init() })
}

/// `src/lib.rs:86-90`: the blur / resize implementation the pyramid is built with.  In the crate the trait has three
/// associated functions; here the arithmetic runs in device kernels and an implementation only names its flavour
/// (`sb200_set_processing`).
pub trait Processing {
    const FLAVOUR: c_int;
}
/// `src/lib.rs:992-1007`: imageproc `gaussian_blur_f32`, image `resize` (Triangle / Nearest) -- the crate's default.
/// Restated from the crates' published algorithms; parity unpinned.
pub struct ImageprocProcessing;
impl Processing for ImageprocProcessing {
    const FLAVOUR: c_int = 1;
}
/// `src/opencv_processing.rs:39-74`: OpenCV `GaussianBlur` / `resize` -- the flavour the crate's test and snapshots pin.
pub struct OpenCVProcessing;
impl Processing for OpenCVProcessing {
    const FLAVOUR: c_int = 0;
}

/// One context per (thread, device); owns the device arenas.
pub struct Context(*mut Sb200Ctx);

impl Context {
    pub fn new(max_w: u32, max_h: u32) -> Self {
        Self::with_processing::<OpenCVProcessing>(max_w, max_h)
    }
    pub fn with_processing<P: Processing>(max_w: u32, max_h: u32) -> Self {
        let mut p = std::ptr::null_mut();
        let st = unsafe { sb200_create(0, max_w, max_h, 1, 0, &mut p) };
        assert!(st == 0, "sb200_create failed with status {st} (no CUDA device? there is no CPU fallback)");
        let ctx = Context(p);
        ctx.check(unsafe { sb200_set_processing(ctx.0, P::FLAVOUR) });
        ctx
    }
    fn check(&self, st: c_int) {
        if st != 0 {
            let msg = unsafe { std::ffi::CStr::from_ptr(sb200_last_error(self.0)) };
            panic!("sift_b200 status {st}: {}", msg.to_string_lossy()); // the crate panics on errors too
        }
    }
    fn take(r: &Sb200Result) -> SiftResult {
        let n = r.n as usize;
        let kps = if n == 0 { Vec::new() } else { unsafe { std::slice::from_raw_parts(r.keypoints, n) }.to_vec() };
        let desc = if n == 0 { Vec::new() } else { unsafe { std::slice::from_raw_parts(r.descriptors, n * 128) }.to_vec() };
        SiftResult { keypoints: kps, descriptors: Array2::from_shape_vec((n, 128), desc).unwrap() }
    }
    pub fn sift(&mut self, img: &GrayImage, features_limit: Option<usize>) -> SiftResult {
        let mut r = std::mem::MaybeUninit::<Sb200Result>::zeroed();
        let st = unsafe {
            sb200_extract(self.0, img.as_raw().as_ptr(), img.width(), img.height(), img.width(),
                          features_limit.map_or(-1, |l| l as i64), r.as_mut_ptr())
        };
        self.check(st);
        Self::take(unsafe { &r.assume_init() })
    }
}

impl Drop for Context {
    fn drop(&mut self) {
        unsafe { sb200_destroy(self.0) }
    }
}

/// `src/lib.rs:124-128`: the Gaussian scale space and the DoG stacks of one image, octave by octave, as
/// `(6, h, w)` / `(5, h, w)` arrays.  The arrays are host copies of the device-resident pyramid, which stays in the
/// context for `sift_with_precomputed`.
pub struct PrecomputedImages {
    ctx: Context,
    pub scale_space: Vec<Array3<f32>>,
    pub dog: Vec<Array3<f32>>,
    pub n_octaves: usize,
}

fn download_stack(ctx: &Context, octave: u32, layers: usize, w: usize, h: usize,
                  f: unsafe extern "C" fn(*mut Sb200Ctx, u32, u32, *mut f32) -> c_int) -> Array3<f32> {
    let mut buf = vec![0f32; layers * h * w];
    for l in 0..layers {
        ctx.check(unsafe { f(ctx.0, octave, l as u32, buf[l * h * w..].as_mut_ptr()) });
    }
    Array3::from_shape_vec((layers, h, w), buf).unwrap()
}

/// src/lib.rs:71-73: `sift_with_processing::<ImageprocProcessing>(img, features_limit)`, as in the crate.
pub fn sift(img: &GrayImage, features_limit: Option<usize>) -> SiftResult {
    sift_with_processing::<ImageprocProcessing>(img, features_limit)
}

/// src/lib.rs:76-81
pub fn sift_with_processing<P: Processing>(img: &GrayImage, features_limit: Option<usize>) -> SiftResult {
    Context::with_processing::<P>(img.width(), img.height()).sift(img, features_limit)
}

/// src/lib.rs:131-143
pub fn precompute_images<P: Processing>(img: &GrayImage) -> PrecomputedImages {
    let ctx = Context::with_processing::<P>(img.width(), img.height());
    let st = unsafe { sb200_precompute(ctx.0, img.as_raw().as_ptr(), img.width(), img.height(), img.width()) };
    ctx.check(st);
    let (mut n, mut ws, mut hs) = (0u32, [0u32; 16], [0u32; 16]);
    ctx.check(unsafe { sb200_pyramid_info(ctx.0, &mut n, ws.as_mut_ptr(), hs.as_mut_ptr(), 16) });
    let mut scale_space = Vec::with_capacity(n as usize);
    let mut dog = Vec::with_capacity(n as usize);
    for o in 0..n {
        let (w, h) = (ws[o as usize] as usize, hs[o as usize] as usize);
        scale_space.push(download_stack(&ctx, o, 6, w, h, sb200_pyramid_layer));
        dog.push(download_stack(&ctx, o, 5, w, h, sb200_pyramid_dog));
    }
    PrecomputedImages { ctx, scale_space, dog, n_octaves: n as usize }
}

/// src/lib.rs:147
pub fn sift_with_precomputed(pre: &PrecomputedImages, features_limit: Option<usize>) -> SiftResult {
    let mut r = std::mem::MaybeUninit::<Sb200Result>::zeroed();
    let st = unsafe { sb200_extract_precomputed(pre.ctx.0, features_limit.map_or(-1, |l| l as i64), r.as_mut_ptr()) };
    pre.ctx.check(st);
    Context::take(unsafe { &r.assume_init() })
}

/// src/lib.rs:785
pub fn compute_descriptor(img: &ArrayView2<f32>, x: f32, y: f32, scale: f32, orientation: f32) -> impl IntoIterator<Item = u8> {
    let (h, w) = img.dim();
    let data = img.as_standard_layout();
    let ctx = Context::new(w as u32, h as u32);
    let k = Sb200DescIn { x, y, scale, orientation };
    let mut out = vec![0u8; 128];
    let st = unsafe { sb200_compute_descriptors(ctx.0, data.as_ptr(), w as u32, h as u32, w as u32, &k, 1, out.as_mut_ptr()) };
    ctx.check(st);
    out
}
