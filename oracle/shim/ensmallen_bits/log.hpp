// oracle/shim: placeholder for an ensmallen header the reference includes (core_private.cpp:8-10); nothing from it is used.
#pragma once
