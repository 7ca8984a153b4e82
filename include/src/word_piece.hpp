#pragma once
#include "../word_piece.hpp"
