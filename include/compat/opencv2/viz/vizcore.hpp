#pragma once
/* visualisation is out of scope */
