"""Seeded synthetic inputs shared by tests/, bench.py and __graft_entry__.smoke(): textures and frame pairs
(textures.py), BASELINE-shaped camera + IMU sequences (sequences.py) and sliding-window BA problems (ba_problems.py).
Neutral data generators: nothing here imports the product (flvis_b200/) or the checker (oracle/)."""
