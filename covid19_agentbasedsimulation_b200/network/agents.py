"""Attribute bags of the routine model -- mirror of covid_abs/network/agents.py (Person, House, Business).

The behaviour methods of the reference (move_*, update, supply, demand, accounting ...) run on the device
(csrc/graph_kernels.cuh); these classes only carry the state `get_population()` / graphics read back.
"""
from enum import Enum

from covid19_agentbasedsimulation_b200.agents import Agent, AgentType


class EconomicalStatus(Enum):   # network/agents.py:9-11
    Active = 1
    Inactive = 0


class Business(Agent):          # network/agents.py:14-31
    def __init__(self, **kwargs):
        super(Business, self).__init__(**kwargs)
        self.employees = []
        self.num_employees = kwargs.get("num_employees", 0)
        self.incomes = kwargs.get("incomes", 0.0)
        self.expenses = kwargs.get("expenses", 0.0)
        self.stocks = kwargs.get("stocks", 10)
        self.sales = kwargs.get("sales", 0)
        self.open = True
        self.type = kwargs.get("type", AgentType.Business)
        self.price = kwargs.get("price", (self.social_stratum + 1) * 12.0)
        self.fixed_expenses = kwargs.get('fixed_expenses', 0.0)


class House(Agent):             # network/agents.py:171-184
    def __init__(self, **kwargs):
        super(House, self).__init__(**kwargs)
        self.homemates = []
        self.type = AgentType.House
        self.size = kwargs.get("size", 0)
        self.incomes = kwargs.get("incomes", 0)
        self.expenses = kwargs.get("expenses", 0)
        self.fixed_expenses = kwargs.get('fixed_expenses', 0.0)


class Person(Agent):            # network/agents.py:256-277
    def __init__(self, **kwargs):
        super(Person, self).__init__(**kwargs)
        self.employer = kwargs.get("employer", None)
        self.house = kwargs.get("house", None)
        self.type = AgentType.Person
        self.economical_status = EconomicalStatus.Inactive
        self.incomes = kwargs.get("income", 0.0)
        self.expenses = kwargs.get("expense", 0.0)
        if self.age > 16 and self.age <= 65:
            self.economical_status = EconomicalStatus.Active

    def is_unemployed(self):
        return self.employer is None and self.economical_status == EconomicalStatus.Active

    def is_homeless(self):
        return self.house is None
