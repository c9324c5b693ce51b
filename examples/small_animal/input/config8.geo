number of panels
8
rotation axis (global frame)
0 0 1
rotation step between panels (degrees, counter-clockwise positive)
45
material id and density (g/cm3): crystal, then gap
7          7.4
0        0.001025
-------------------------------------------------
index of the first panel
0
panel size x y z (cm)
2 15.91 22.99
module size x y z (cm)
2 1.75 1.75
module gap x y z (cm)
2 0.02 0.02
crystal size x y z (cm)
2 0.21 0.21
crystal gap x y z (cm)
2 0.01 0.01
growth direction along local x y z
-1  1 1
centre of the panel face looking at the phantom (cm)
0  -22.5   0
local x axis in the global frame
0  1   0
local y axis in the global frame
1  0   0
local z axis in the global frame
0   0  -1
-------------------------------------------------
