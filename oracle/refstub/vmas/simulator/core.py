"""vmas.simulator.core: World / Agent / Landmark / Sphere, restated on the CPU (TEST INFRASTRUCTURE).

The physics is the vmas 1.4.0 world step as SURVEY.md Appendix A specifies it, expressed with the SAME torch op
sequence as the pinned oracle (oracle/swarm_oracle.py OracleWorld.step, which reproduces the reference's shipped
trajectories bit for bit): forces start from the decoded action, sphere-sphere contact forces for every collidable pair
a < b in entity order (landmarks first, then agents) that passes vmas' ``collides`` pre-filter, drag, semi-implicit
Euler.  Generic over the entity set, so that any of the reference's four scenario files can build its world on it.
"""
import torch

from oracle import swarm_oracle as so

from .utils import Color


class Shape:
    pass


class Sphere(Shape):
    def __init__(self, radius: float = 0.05):
        assert radius > 0
        self.radius = radius

    def circumscribed_radius(self):
        return self.radius


class EntityState:
    def __init__(self):
        self.pos = None
        self.vel = None


class Action:
    def __init__(self):
        self.u = None


class Entity:
    def __init__(self, name, movable=False, rotatable=False, collide=True, mass=1.0, shape=None, color=Color.GRAY,
                 **_ignored):
        self.name = name
        self.movable, self.rotatable, self.collide = movable, rotatable, collide
        self.mass = mass
        self.shape = shape if shape is not None else Sphere()
        self.color = color
        self.state = EntityState()
        self.batch_dim = None
        self.device = None

    def _spawn(self, batch_dim, device):
        self.batch_dim, self.device = batch_dim, device
        self.state.pos = torch.zeros(batch_dim, 2)
        self.state.vel = torch.zeros(batch_dim, 2)

    def _set(self, name, new, batch_index):
        # vmas Entity._set_state_property
        new = torch.as_tensor(new, dtype=torch.float32)
        if batch_index is None:
            if new.dim() > 1 and new.shape[0] == self.batch_dim:
                setattr(self.state, name, new.clone())
            else:
                setattr(self.state, name, new.repeat(self.batch_dim, 1))
        else:
            value = getattr(self.state, name)
            value[batch_index] = new

    def set_pos(self, pos, batch_index):
        self._set("pos", pos, batch_index)

    def set_vel(self, vel, batch_index):
        self._set("vel", vel, batch_index)


class Landmark(Entity):
    def __init__(self, name, shape=None, movable=False, rotatable=False, collide=True, color=Color.GRAY, **kw):
        super().__init__(name, movable=movable, rotatable=rotatable, collide=collide, shape=shape, color=color, **kw)


class Agent(Entity):
    def __init__(self, name, shape=None, movable=True, rotatable=True, collide=True, color=Color.BLUE,
                 render_action=False, u_range=1.0, u_multiplier=1.0, **kw):
        super().__init__(name, movable=movable, rotatable=rotatable, collide=collide, shape=shape, color=color, **kw)
        assert u_range == 1.0 and u_multiplier == 1.0
        self.action = Action()


class World:
    def __init__(self, batch_dim, device, dt=so.DT, substeps=1, drag=so.DRAG, collision_force=so.COLLISION_FORCE,
                 contact_margin=so.CONTACT_MARGIN, **_ignored):
        assert substeps == 1 and (dt, drag, collision_force, contact_margin) == (so.DT, so.DRAG, so.COLLISION_FORCE,
                                                                                 so.CONTACT_MARGIN), \
            "the restated step covers the vmas defaults the reference runs with"
        self.batch_dim = batch_dim
        self.device = torch.device(device)
        self._agents, self._landmarks = [], []

    @property
    def agents(self):
        return self._agents

    @property
    def landmarks(self):
        return self._landmarks

    @property
    def entities(self):
        return self._landmarks + self._agents

    def add_agent(self, agent):
        agent._spawn(self.batch_dim, self.device)
        self._agents.append(agent)

    def add_landmark(self, landmark):
        landmark._spawn(self.batch_dim, self.device)
        self._landmarks.append(landmark)

    def reset(self, env_index):
        for e in self.entities:
            if env_index is None:
                e.state.pos = torch.zeros(self.batch_dim, 2)
                e.state.vel = torch.zeros(self.batch_dim, 2)
            else:
                e.state.pos[env_index] = 0.0
                e.state.vel[env_index] = 0.0

    def get_distance(self, a, b):
        # sphere-sphere: get_distance_from_point subtracts r_a, the caller then r_b
        assert isinstance(a.shape, Sphere) and isinstance(b.shape, Sphere)
        return (torch.linalg.vector_norm(a.state.pos - b.state.pos, dim=-1) - a.shape.radius) - b.shape.radius

    def _collides(self, a, b):
        if not a.collide or not b.collide or a is b:
            return False
        if not a.movable and not a.rotatable and not b.movable and not b.rotatable:
            return False
        d = torch.linalg.vector_norm(a.state.pos - b.state.pos, dim=-1)
        return bool((d <= a.shape.circumscribed_radius() + b.shape.circumscribed_radius()).any())

    def step(self):
        ents = self.entities
        forces = {}
        for ag in self._agents:
            forces[ag] = torch.zeros(self.batch_dim, 2) + ag.action.u
        pairs = [(a, b) for ia, a in enumerate(ents) for ib, b in enumerate(ents) if ib > ia and self._collides(a, b)]
        if pairs:
            for a, b in pairs:
                assert a.shape.radius == so.SPHERE_RADIUS and b.shape.radius == so.SPHERE_RADIUS
            pos_a = torch.stack([a.state.pos for a, _ in pairs], dim=-2)
            pos_b = torch.stack([b.state.pos for _, b in pairs], dim=-2)
            force_a = so.constraint_force(pos_a, pos_b)
            force_b = -force_a
            for p, (a, b) in enumerate(pairs):
                if a.movable:
                    forces[a] = forces[a] + force_a[:, p]
                if b.movable:
                    forces[b] = forces[b] + force_b[:, p]
        for ag in self._agents:
            ag.state.vel = ag.state.vel * (1 - so.DRAG)
            accel = forces[ag] / ag.mass
            ag.state.vel = ag.state.vel + accel * so.DT
            ag.state.pos = ag.state.pos + ag.state.vel * so.DT
