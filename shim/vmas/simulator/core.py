from swarm_b200.scenarios import Agent, Box, Entity, Landmark, Line, Shape, Sphere, World   # noqa: F401
