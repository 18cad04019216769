"""Agent record and its enumerations -- host-side mirror of covid_abs/agents.py.

Same names, values and declaration order as the reference (agents.py:9-37): the declaration order of
`Status` defines the key order of `get_statistics` (abs.py:300-302).  On the device the population is
held as structure of arrays; `Agent` objects are materialised lazily by `Simulation.get_population`.
"""
from enum import Enum
import uuid


class Status(Enum):
    """Agent status, following the SIR model (agents.py:9-16)."""
    Susceptible = 's'
    Infected = 'i'
    Recovered_Immune = 'c'
    Death = 'm'


class InfectionSeverity(Enum):
    """Severity of the Infected agents (agents.py:19-26)."""
    Exposed = 'e'
    Asymptomatic = 'a'
    Hospitalization = 'h'
    Severe = 'g'


class AgentType(Enum):
    """Type of the agent / node of the graph (agents.py:29-37)."""
    Person = 'p'
    Business = 'b'
    House = 'h'
    Government = 'g'
    Healthcare = 'c'


# device encodings (include/covidabs.h)
STATUS_LIST = [Status.Susceptible, Status.Infected, Status.Recovered_Immune, Status.Death]
STATUS_CODE = {s: i for i, s in enumerate(STATUS_LIST)}
SEVERITY_LIST = [InfectionSeverity.Exposed, InfectionSeverity.Asymptomatic, InfectionSeverity.Hospitalization,
                 InfectionSeverity.Severe]
SEVERITY_CODE = {s: i for i, s in enumerate(SEVERITY_LIST)}


class Agent(object):
    """Container of an agent's attributes (agents.py:40-64)."""

    def __init__(self, **kwargs):
        self.id = kwargs.get('id', int(uuid.uuid4()))
        self.x = kwargs.get('x', 0)
        self.y = kwargs.get('y', 0)
        self.status = kwargs.get('status', Status.Susceptible)
        self.infected_status = kwargs.get('infected_status', InfectionSeverity.Asymptomatic)
        self.infected_time = kwargs.get('infected_time', 0)
        self.age = kwargs.get('age', 0)
        self.social_stratum = kwargs.get('social_stratum', 0)
        self.wealth = kwargs.get('wealth', 0.0)
        self.type = AgentType.Person
        self.environment = kwargs.get('environment', None)

    def get_description(self):
        """Simplified description of the health status (agents.py:66-75)."""
        if self.status == Status.Infected:
            return "{}({})".format(self.status.name, self.infected_status.name)
        return self.status.name

    def __str__(self):
        return str(self.status.name)
