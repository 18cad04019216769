"""Agent record and its enumerations -- host-side mirror of covid_abs/agents.py.

Names, values and declaration order are the reference's (agents.py:9-37): scenario dictionaries are keyed by
`Status` members, statistics keys are the member names, and the declaration order of `Status` is the key order of
`get_statistics` (abs.py:300-302).  On the device a population is structure of arrays; `Agent` objects are only
materialised on request (`Simulation.get_population`).
"""
from enum import Enum
import uuid

# member name -> one-letter value, in the reference's declaration order
Status = Enum('Status', [('Susceptible', 's'), ('Infected', 'i'), ('Recovered_Immune', 'c'), ('Death', 'm')])
Status.__doc__ = "Health status of an agent, SIR plus death (agents.py:9-16)."
InfectionSeverity = Enum('InfectionSeverity', [('Exposed', 'e'), ('Asymptomatic', 'a'), ('Hospitalization', 'h'),
                                               ('Severe', 'g')])
InfectionSeverity.__doc__ = "Severity of an infection (agents.py:19-26)."
AgentType = Enum('AgentType', [('Person', 'p'), ('Business', 'b'), ('House', 'h'), ('Government', 'g'),
                               ('Healthcare', 'c')])
AgentType.__doc__ = "Kind of node in the routine model (agents.py:29-37)."

# device encodings (include/covidabs.h): position in the declaration order
STATUS_LIST = list(Status)
STATUS_CODE = {member: code for code, member in enumerate(STATUS_LIST)}
SEVERITY_LIST = list(InfectionSeverity)
SEVERITY_CODE = {member: code for code, member in enumerate(SEVERITY_LIST)}

# attribute -> default of an Agent (agents.py:44-64)
_AGENT_FIELDS = (('x', 0), ('y', 0), ('status', Status.Susceptible), ('infected_status', InfectionSeverity.Asymptomatic),
                 ('infected_time', 0), ('age', 0), ('social_stratum', 0), ('wealth', 0.0), ('environment', None))


class Agent(object):
    """Attribute bag of one agent (agents.py:40-64).  `id` defaults to a random 128-bit integer like the reference's;
    agents materialised from the device carry their index in the population list instead."""

    def __init__(self, **kwargs):
        self.id = kwargs['id'] if 'id' in kwargs else int(uuid.uuid4())
        for name, default in _AGENT_FIELDS:
            setattr(self, name, kwargs.get(name, default))
        self.type = AgentType.Person

    def get_description(self):
        """Status name, with the severity appended for the Infected (agents.py:66-75)."""
        name = self.status.name
        return "{}({})".format(name, self.infected_status.name) if self.status == Status.Infected else name

    def __str__(self):
        return str(self.status.name)
