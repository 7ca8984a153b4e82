#pragma once
#include "../word_piece_utils.hpp"
