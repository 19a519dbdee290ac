//! examples/benchmark_gpu.rs — the teapot benchmark of `examples/benchmark.rs` on the batched GPU path.
//!
//!     RTBVH_B200_LIB_DIR=/path/to/repo/rtbvh_b200 cargo run --release --example benchmark_gpu
//!
//! Same scene, same camera, same 100 frames of 1000x1000 rays; where the reference times
//! `for (triangle, ray) in bvh.iter(ray, triangles) { triangle.intersect(ray); }` over rayon chunks
//! (`examples/benchmark.rs:25-31`, `:55-61`), this one hands the whole frame to `GpuScene` and then replays a sample
//! through the host iterators to show that `ray.t` is bit-identical.  Not compiled here (no cargo in the image).
use glam::*;
use rtbvh::*;
use shared::*;
use std::num::NonZeroUsize;

const WIDTH: usize = 1000;
const HEIGHT: usize = 1000;
const FRAMES: usize = 100;

/// `v x y z` / `f a b c ...` only (what teapot.obj holds); faces are fanned, triangle order = file order.
fn load_obj(text: &str) -> Vec<Triangle> {
    let mut points: Vec<Vec4> = Vec::new();
    let mut triangles = Vec::new();
    for line in text.lines() {
        let mut words = line.split_whitespace();
        match words.next() {
            Some("v") => {
                let c: Vec<f32> = words.take(3).map(|w| w.parse().expect("vertex coordinate")).collect();
                points.push(vec4(c[0], c[1], c[2], 1.0));
            }
            Some("f") => {
                let corner = |w: &str| -> usize {
                    let i: i64 = w.split('/').next().unwrap().parse().expect("face index");
                    if i < 0 { (points.len() as i64 + i) as usize } else { i as usize - 1 }
                };
                let ids: Vec<usize> = words.map(corner).collect();
                for k in 1..ids.len().saturating_sub(1) {
                    triangles.push(Triangle::new(points[ids[0]], points[ids[k]], points[ids[k + 1]]));
                }
            }
            _ => {}
        }
    }
    triangles
}

/// The camera of `examples/benchmark.rs:74-99` (including its doubly converted field of view).
fn benchmark_camera() -> CameraView3D {
    let fov = 90_f32.to_radians();
    let half = (fov * 0.5 / (180.0 / std::f32::consts::PI)).tan();
    let aspect_ratio = WIDTH as f32 / HEIGHT as f32;
    let (pos, forward, up) = (vec3(0.0, 1.5, -100.0), Vec3::Z, Vec3::Y);
    let right = forward.cross(up);
    let center = pos + forward;
    let p1 = center - half * right * aspect_ratio + half * up;
    let p2 = center + half * right * aspect_ratio + half * up;
    let p3 = center - half * right * aspect_ratio - half * up;
    CameraView3D {
        pos,
        right: p2 - p1,
        up: p3 - p1,
        p1,
        direction: forward,
        inv_width: 1.0 / WIDTH as f32,
        inv_height: 1.0 / HEIGHT as f32,
        aspect_ratio,
        fov,
    }
}

fn main() -> Result<(), Box<dyn std::error::Error>> {
    assert!(device_count() > 0, "no CUDA device: this crate has no CPU fallback");
    let triangles = load_obj(include_str!("../objects/teapot.obj"));
    let camera = benchmark_camera();

    // Builder / Mbvh::construct are the crate's own API; with the patch applied both run on the GPU.
    let timer = Timer::default();
    let bvh = Builder { aabbs: None, primitives: &triangles, primitives_per_leaf: NonZeroUsize::new(1) }.construct_binned_sah()?;
    println!("binned SAH build ({} triangles): {:.3} ms, validate = {}", triangles.len(), timer.elapsed_in_millis(), bvh.validate(triangles.len()));
    let timer = Timer::default();
    let mbvh = Mbvh::construct(&bvh);
    println!("Mbvh::construct: {:.3} ms, {} quad nodes", timer.elapsed_in_millis(), mbvh.quad_nodes().len());
    let (device_ms, total_ms, _) = last_build_stats()?;
    println!("  last GPU job: {:.3} ms of kernels, {:.3} ms including copies", device_ms, total_ms);

    let scene = GpuScene::new(Some(&bvh), Some(&mbvh), &triangles)?;
    let frame: Vec<Ray> = (0..HEIGHT)
        .flat_map(|y| (0..WIDTH).map(move |x| (x, y)))
        .map(|(x, y)| camera.generate_ray(x as u32, y as u32))
        .collect();
    let packet_frame: Vec<RayPacket4> = (0..HEIGHT)
        .flat_map(|y| (0..WIDTH).step_by(4).map(move |x| (x as u32, y as u32)))
        .map(|(x, y)| camera.generate_ray4([x, x + 1, x + 2, x + 3], [y; 4]))
        .collect();

    for (name, tree) in [("Bvh", Tree::Bvh), ("Mbvh", Tree::Mbvh)] {
        let timer = Timer::default();
        let mut hit_count = 0usize;
        for _ in 0..FRAMES {
            let mut rays = frame.clone();
            let hits = scene.intersect(tree, &mut rays)?;
            hit_count += hits.iter().filter(|h| h.prim.is_some()).count();
        }
        let ms = timer.elapsed_in_millis();
        println!("{:4} single rays:  {} rays in {:.1} ms = {:.1} Mrays/s ({} hits)", name, FRAMES * frame.len(), ms, (FRAMES * frame.len()) as f32 / ms / 1000.0, hit_count);

        let timer = Timer::default();
        for _ in 0..FRAMES {
            let mut packets = packet_frame.clone();
            scene.intersect_packets(tree, &mut packets, 1e-4)?;
        }
        let ms = timer.elapsed_in_millis();
        println!("{:4} packets of 4:  {} rays in {:.1} ms = {:.1} Mrays/s", name, FRAMES * frame.len(), ms, (FRAMES * frame.len()) as f32 / ms / 1000.0);
    }

    // Parity: every 97th ray through the reference's own loop on the host.
    let mut rays = frame.clone();
    let hits = scene.intersect(Tree::Mbvh, &mut rays)?;
    let mut checked = 0;
    for k in (0..frame.len()).step_by(97) {
        let mut ray = frame[k];
        let mut prim = None;
        for (id, r) in mbvh.traverse_iter_indices(&mut ray) {
            if triangles[id as usize].intersect(r) {
                prim = Some(id);
            }
        }
        assert_eq!(ray.t.to_bits(), rays[k].t.to_bits(), "ray {}: t differs", k);
        assert_eq!(prim.is_some(), hits[k].prim.is_some(), "ray {}: hit / miss differs", k);
        checked += 1;
    }
    println!("parity: {} sampled rays bit-identical to the host iterator loop", checked);
    Ok(())
}
