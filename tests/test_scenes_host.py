"""Host generators that double as parity references for the device-side seeding / level-set rasterisation
(SURVEY §8f row 4): pure numpy, no GPU."""
import numpy as np

from taichi_mpm_b200 import capi, scenes


def test_hashed_lattice_without_jitter_is_the_reference_lattice():
    # scenes.lattice_block is pinned to the reference's add_particles(benchmark=...) (tests/test_oracle_ref_transfer.py)
    res, lo, hi = 32, (9, 10, 11), (14, 13, 17)
    ids, x, mass, vol = scenes.lattice_block_hashed(res, lo, hi, jitter=0.0)
    xr, mr, vr = scenes.lattice_block(res, lo, hi)
    assert np.array_equal(ids, np.arange(len(xr), dtype=np.uint32))
    assert np.abs(x - xr).max() <= 1e-7 and np.array_equal(mass, mr) and np.array_equal(vol, vr)


def test_hashed_lattice_is_partition_independent_and_jitter_is_bounded():
    res, lo, hi = 32, (9, 10, 8), (14, 13, 24)
    ids, x, _, _ = scenes.lattice_block_hashed(res, lo, hi, jitter=0.1, seed=3)
    parts = [scenes.lattice_block_hashed(res, lo, hi, jitter=0.1, seed=3, z_cells=zc) for zc in ((0, 5), (5, 11), (11, 16))]
    allids = np.concatenate([p[0] for p in parts])
    o = np.argsort(allids, kind="stable")
    assert np.array_equal(allids[o], ids) and np.array_equal(np.concatenate([p[1] for p in parts])[o], x)
    _, x0, _, _ = scenes.lattice_block_hashed(res, lo, hi, jitter=0.0)
    d = np.abs(x - x0) * res
    assert 0.05 < d.max() <= 0.1 + 1e-6 and abs(d.mean() - 0.05) < 0.005       # uniform in [-0.1, 0.1)
    other = scenes.lattice_block_hashed(res, lo, hi, jitter=0.1, seed=4)[1]
    assert np.abs(other - x).max() > 0.01 / res                                  # the seed matters


def test_shapes_twin_planes_equal_the_plane_rasteriser_and_solids_are_signed_distances():
    res = 24
    planes = np.array([[0.0, 1.0, 0.0, -9.5], [0.6, 0.8, 0.0, -3.0]], np.float32)
    a = scenes.planes_sdf(res, planes)
    b = scenes.shapes_sdf(res, [(capi.SHAPE_PLANE, False, list(p)) for p in planes])
    assert np.array_equal(a, b)
    s = scenes.shapes_sdf(res, [(capi.SHAPE_SPHERE, False, [12.0, 12.0, 12.0, 5.0])])
    X = np.stack(np.meshgrid(*[np.arange(res + 1, dtype=np.float64)] * 3, indexing="ij"), -1)
    r = np.linalg.norm(X - 12.0, axis=-1)
    assert np.abs(s[..., 3] - (r - 5.0)).max() < 1e-5
    assert np.abs(np.linalg.norm(s[..., :3], axis=-1) - 1.0).max() < 1e-5
    c = scenes.shapes_sdf(res, [(capi.SHAPE_CUBOID, True, [4.0, 6.0, 5.0, 20.0, 18.0, 19.0])])      # a container
    assert c[12, 12, 12, 3] > 0 and abs(c[12, 12, 12, 3] - 6.0) < 1e-6           # free space inside, 6 nodes from the nearest wall (y)
    assert abs(c[12, 3, 12, 3] + 3.0) < 1e-6 and np.allclose(c[12, 3, 12, :3], [0, 1, 0])   # 3 nodes into the floor, normal points back in
    assert abs(c[2, 4, 12, 3] + np.hypot(2.0, 2.0)) < 1e-5                       # an edge: Euclidean distance to the box


def test_host_rigid_body_levelset_collision_removes_the_approach_velocity():
    # MPM<dim>::rigid_body_levelset_collision (src/mpm_rigid_body.cpp:346-381) on the host-side body: a box hitting a floor with one
    # corner — the contact point stops approaching (restitution 0), the body starts to rotate, friction bounds the tangential impulse
    import numpy as np
    from taichi_mpm_b200 import rigid, scenes
    b = rigid.HostRigidBody(scenes.box_mesh((0.1, 0.05, 0.08)), density=1000.0, position=(0.5, 0.3, 0.5), euler_deg=(0, 0, 20.0),
                            velocity=(0.5, -2.0, 0.0), frictions=(0.3, 0.3))
    corners = b.position + np.array([[sx * 0.1, sy * 0.05, sz * 0.08] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)]) @ b.rotation.T
    floor_y = corners[:, 1].min() + 0.004                         # the lowest corners are 4 mm inside
    phi = corners[:, 1] - floor_y
    g = np.tile([0.0, 1.0, 0.0], (8, 1))
    inside = phi < 0
    assert 1 <= inside.sum() <= 2
    m0 = b.mass * b.velocity
    hits = b.levelset_collision(corners, phi, g)
    assert hits >= 1
    for p in corners[inside]:
        # one pass over the samples, as the reference loops: each impulse zeroes its own contact's approach velocity, the friction
        # impulse and the next contact then disturb it a little
        assert b.velocity_at(p)[1] >= -0.5                        # from -2.0 .. -2.3 before
    assert np.abs(b.angular_velocity).max() > 0.1                 # an off-centre contact spins the box
    dj = b.mass * b.velocity - m0
    assert dj[1] > 0 and np.hypot(dj[0], dj[2]) <= 0.3 * dj[1] * (1 + 1e-9) * hits   # Coulomb bound on the tangential impulse
    k = rigid.HostRigidBody(scenes.box_mesh((0.1, 0.05, 0.08)), scripted_position=lambda t: (0.5, 0.3, 0.5), scripted_rotation=lambda t: (0, 0, 0))
    assert k.levelset_collision(corners, phi, g) == 0             # scripted bodies take no impulse
