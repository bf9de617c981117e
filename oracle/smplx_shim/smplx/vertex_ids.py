"""ORACLE (test infrastructure): SMPL-H vertex ids used by VertexJointSelector."""
vertex_ids = {
    'smplh': {
        'nose': 332, 'reye': 6260, 'leye': 2800, 'rear': 4071, 'lear': 583,
        'rthumb': 6191, 'rindex': 5782, 'rmiddle': 5905, 'rring': 6016, 'rpinky': 6133,
        'lthumb': 2746, 'lindex': 2319, 'lmiddle': 2445, 'lring': 2556, 'lpinky': 2673,
        'LBigToe': 3216, 'LSmallToe': 3226, 'LHeel': 3387,
        'RBigToe': 6617, 'RSmallToe': 6624, 'RHeel': 6787,
    }
}
