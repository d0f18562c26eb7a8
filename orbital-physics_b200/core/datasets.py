"""J2000 Keplerian elements of the solar system (static table, host only).

Same bodies, order and numbers as the reference's core/datasets.py:13-56
(values: JPL approximate planetary elements / ssd.jpl.nasa.gov satellite
elements); laid out as tables instead of constructor calls.
"""
from __future__ import annotations

from core.body import Body, System
from core.constants import J2000_JD, STANDARD
from core.units import AU, Degrees, Kilograms, Meters

G = STANDARD.G
EPOCH = J2000_JD   # https://en.wikipedia.org/wiki/Epoch_(astronomy)#J2000

# name, mass [kg], radius [m], a [AU], e, I, L, long_peri, long_node  (angles in degrees)
_HELIOCENTRIC = (
    ("Mercury", 3.3011e23, 2.4397e6, 0.38709927, 0.20563593, 7.00497902, 252.25032350, 77.45779628, 48.33076593),
    ("Venus", 4.8675e24, 6.0518e6, 0.72333566, 0.00677672, 3.39467605, 181.97909950, 131.60246718, 76.67984255),
    ("Earth", 5.9722e24, 6.371e6, 1.00000261, 0.01671123, -0.00001531, 100.46457166, 102.93768193, 0.0),
    ("Mars", 6.4171e23, 3.3895e6, 1.52371034, 0.09339410, 1.84969142, -4.55343205, -23.94362959, 49.55953891),
    ("Jupiter", 1.8982e27, 6.9911e7, 5.20288700, 0.04838624, 1.30439695, 34.39644051, 14.72847983, 100.47390909),
    ("Saturn", 5.6834e26, 5.8232e7, 9.53667594, 0.05386179, 2.48599187, 49.95424423, 92.59887831, 113.66242448),
    ("Uranus", 8.6810e25, 2.5362e7, 19.18916464, 0.04725744, 0.77263783, 313.23810451, 170.95427630, 74.01692503),
    ("Neptune", 1.02413e26, 2.4622e7, 30.06992276, 0.00859048, 1.77004347, -55.12002969, 44.96476227, 131.78422574),
    ("Pluto", 13024.6e18, 1188300, 39.5886, 0.2518, 17.1477, 38.68366, 113.709, 110.292),
    ("Ceres", 938.416e18, 469700, 2.766051, 0.0794, 10.588, 188.70268, 73.2734, 80.2522),
    ("Eris", 16600e18, 1163000, 68.0506, 0.435675, 43.821, 211.032, 150.714, 36.0460),
    ("20000 Varuna", 3.698e20, 334000, 43.1374, 0.053565, 17.1395, 114.900, 272.579, 97.21338),
    ("Makemake", 3100e18, 714000, 45.4494, 0.16194, 29.03386, 168.8258, 296.95, 79.259),
    ("28978 Ixion", 3e20, 355000, 39.3745, 0.2449, 19.6745, 293.546, 300.585, 71.099),
)

# parent, name, mass [kg], radius [m], a, e, I, arg_peri, M, long_node  (https://ssd.jpl.nasa.gov/sats/elem/)
_SATELLITES = (
    ("Earth", "Luna", 7.346e22, 1.7371e6, AU(0.00257), 0.0549, 5.16, 318.15, 135.27, 125.08),
    ("Jupiter", "Io", 8.93e22, 1_821_600, Meters(421_800_000), 0.004, 0., 49.1, 330.9, 0.),
    ("Jupiter", "Europa", 4.8e22, 1_560_800, Meters(671_100_000), 0.009, 0.5, 45.0, 345.4, 184.0),
    ("Jupiter", "Ganymede", 1.4819e23, 2_634_100, Meters(1_070_400_000), 0.001, 0.2, 198.3, 324.8, 58.5),
    ("Jupiter", "Callisto", 1.08e23, 1_560_800, Meters(1_882_700_000), 0.007, 0.3, 43.8, 87.4, 309.1),
    ("Saturn", "Titan", 1.345e23, 2_575_000, Meters(1_221_900_000), 0.029, 0.35, 78.3, 11.7, 78.6),
    ("Saturn", "Enceladus", 1.08e20, 252_000, Meters(238_400_000), 0.005, 0.0, 119.5, 57.0, 0.0),
    ("Saturn", "Rhea", 2.31e21, 763_800, Meters(527_200_000), 0.001, 0.3, 44.3, 31.5, 133.7),
    ("Saturn", "Iapetus", 1.805e21, 734_400, Meters(3_561_700_000), 0.028, 7.6, 254.5, 74.8, 86.5),
    ("Neptune", "Triton", 2.14e22, 1_353_400, Meters(354_800_000), 0.0, 157.3, 0.0, 63.0, 178.1),
    ("Uranus", "Titania", 3.455e21, 788_400, Meters(436_298_000), 0.002, 0.1, 184.0, 68.1, 29.5),
)


def solar_system_v2(moons: bool = False, **kwargs) -> System:
    """Sun, 8 planets, 6 dwarf/minor planets (15 bodies); with `moons`, 11 satellites more."""
    sol = Body(parent=None, name="Sol", mass=Kilograms(1.9885e30), radius=Meters(6.9634e8), a=AU(0), e=0,
               I=Degrees(0), L=Degrees(0), long_peri=Degrees(0), long_node=Degrees(0), arg_peri=None, M=None)
    bodies = [sol]
    by_name = {"Sol": sol}
    for name, mass, radius, a, e, inc, L, varpi, node in _HELIOCENTRIC:
        body = Body(parent=sol, name=name, mass=Kilograms(mass), radius=Meters(radius), a=AU(a), e=e,
                    I=Degrees(inc), L=Degrees(L), long_peri=Degrees(varpi), long_node=Degrees(node),
                    M=None, arg_peri=None)
        bodies.append(body)
        by_name[name] = body
    if moons:
        for parent, name, mass, radius, a, e, inc, argp, M, node in _SATELLITES:
            a_au = a.to_au() if isinstance(a, Meters) else a
            bodies.append(Body(parent=by_name[parent], name=name, mass=Kilograms(mass), radius=Meters(radius),
                               a=a_au, e=e, I=Degrees(inc), arg_peri=Degrees(argp), M=Degrees(M),
                               long_node=Degrees(node), long_peri=None, L=None))
    return System(bodies, **kwargs)


solar_system = solar_system_v2   # backwards-compatible alias
