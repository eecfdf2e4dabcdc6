"""CPU: vehicle boxes -> obstacle array, against an independent restatement of the matrix
pipeline of the reference (CudaAgent._create_bb_points / get_matrix / _vehicle_to_world / _trace_route,
carla/cuda_agent.py:127-206) for planar poses."""
import numpy as np

import synthetic_world as sw


def _reference_pipeline(vehicles, ego_xy):
    """cuda_agent.py:127-206 restated with 4x4 matrices (pitch = roll = 0, z = 0)."""
    def matrix(x, y, yaw_deg):                      # get_matrix, :149-172
        cy, sy = np.cos(np.radians(yaw_deg)), np.sin(np.radians(yaw_deg))
        m = np.identity(4)
        m[0, 3], m[1, 3] = x, y
        m[0, 0], m[0, 1], m[1, 0], m[1, 1] = cy, -sy, sy, cy
        return m
    out = []
    for x, y, yaw, ex, ey, ox, oy in vehicles:
        cords = np.zeros((3, 4))                    # _create_bb_points, :127-137
        cords[0] = [ex + 2.5, ey + 2.5, 0, 1]
        cords[1] = [-ex - 2.5, -ey - 2.5, 0, 1]
        cords[2] = [0, 0, 0, 1]
        world = matrix(x, y, yaw) @ matrix(ox, oy, 0.0) @ cords.T       # _vehicle_to_world, :139-147
        world /= world[3]
        dist = np.sqrt((ego_xy[0] - world[0, 2]) ** 2 + (ego_xy[1] - world[1, 2]) ** 2)
        if 0 < dist <= 30:                           # :193
            out.append([world[0, 0], world[1, 0], world[0, 1], world[1, 1]])
    if not out:
        return np.array([[-1, -1, -1, -1]], np.float32), np.array([0], np.int32)
    return np.array(out).astype(np.float32), np.array([len(out)], np.int32)


def test_matches_the_reference_matrix_pipeline():
    rng = np.random.default_rng(4)
    veh = np.stack([rng.uniform(-40, 40, 60), rng.uniform(-40, 40, 60), rng.uniform(-180, 180, 60),
                    rng.uniform(1.5, 3.0, 60), rng.uniform(0.7, 1.2, 60), rng.uniform(-0.3, 0.3, 60),
                    rng.uniform(-0.1, 0.1, 60)], 1)
    ego = (1.0, -2.0)
    obs, num = sw.vehicles_to_obstacles(veh, ego)
    robs, rnum = _reference_pipeline(veh, ego)
    assert num[0] == rnum[0] and 0 < num[0] < 60       # some vehicles are beyond 30 m
    np.testing.assert_allclose(obs, robs, rtol=0, atol=1e-5)


def test_ego_is_dropped_and_empty_gives_the_placeholder():
    ego_vehicle = [[5.0, 5.0, 30.0, 2.0, 1.0, 0.0, 0.0]]
    obs, num = sw.vehicles_to_obstacles(ego_vehicle, (5.0, 5.0))        # distance 0: not an obstacle (:193)
    assert num[0] == 0 and obs.tolist() == [[-1, -1, -1, -1]]
    obs, num = sw.vehicles_to_obstacles(np.zeros((0, 7)), (0.0, 0.0))
    assert num[0] == 0 and obs.shape == (1, 4)


def test_turned_vehicle_quirk_and_the_oriented_extension():
    # a car turned by 45 degrees: the reference's two diagonal corners land on one axis -> a sliver
    veh = [[10.0, 0.0, 45.0, 2.0, 2.0, 0.0, 0.0]]
    obs, num = sw.vehicles_to_obstacles(veh, (0.0, 0.0))
    assert num[0] == 1
    w, h = abs(obs[0, 2] - obs[0, 0]), abs(obs[0, 3] - obs[0, 1])
    assert w < 1e-4 and abs(h - 2 * 4.5 * np.sqrt(2)) < 1e-4
    rect, num = sw.vehicles_to_obstacles(veh, (0.0, 0.0), oriented=True)
    np.testing.assert_allclose(rect[0], [10.0, 0.0, 4.5, 4.5, np.pi / 4], atol=1e-6)
    # axis-aligned car: both forms describe the same box
    veh = [[10.0, 3.0, 0.0, 2.0, 1.0, 0.5, 0.0]]
    obs, _ = sw.vehicles_to_obstacles(veh, (0.0, 0.0))
    rect, _ = sw.vehicles_to_obstacles(veh, (0.0, 0.0), oriented=True)
    assert sorted([obs[0, 0], obs[0, 2]]) == [rect[0, 0] - rect[0, 2], rect[0, 0] + rect[0, 2]]
    assert sorted([obs[0, 1], obs[0, 3]]) == [rect[0, 1] - rect[0, 3], rect[0, 1] + rect[0, 3]]
