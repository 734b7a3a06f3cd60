"""Small helpers shared by the tests."""
from oracle import ba_ref


def oracle_data(p):
    """synthdata.ba_problems.Problem -> oracle.ba_ref.BAData (fresh copies of the state arrays)."""
    return ba_ref.BAData(p.poses.copy(), p.lms.copy(), p.ep, p.el, p.uv, p.K, p.fixed_pose, p.fix_landmarks)
